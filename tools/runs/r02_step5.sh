#!/bin/bash
# 2 GPUs: full gpu tests (incl. push exchange check + model surface), TP=2 bench with the fused exchange
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
for ex in auto push-nomc; do
MIXQ_TP_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_tp2_$ex.json 2> gpurun_out/r02_bench_tp2_$ex.err
echo "bench tp2 $ex rc=$?"; tail -3 gpurun_out/r02_bench_tp2_$ex.err
done
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_tp2_auto.json", "gpurun_out/r02_bench_tp2_push-nomc.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("tp_parity"), d.get("step_breakdown_us"), d["config"].get("exchange"), {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
