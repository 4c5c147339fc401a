#!/bin/bash
# 2 GPUs: tests (W4 on the 2-CTA kernel, exchange v2 with one-shot / two-phase), exchange anatomy, TP=2 bench, C3 bench
mkdir -p gpurun_out
export MIXQ_PEER_TIMEOUT_MS=20000
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/bench_exchange.py > gpurun_out/r02_bench_exchange_2b.json 2> gpurun_out/r02_bench_exchange_2b.err
echo "rc=$?"; cat gpurun_out/r02_bench_exchange_2b.json
for os in 1 0; do
MIXQ_TP_ONE_SHOT=$os timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_tp2_os$os.json 2> gpurun_out/r02_bench_tp2_os$os.err
echo "bench tp2 one_shot=$os rc=$?"; tail -2 gpurun_out/r02_bench_tp2_os$os.err
done
timeout 600 python bench.py --bit 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c3_w4.json 2> gpurun_out/r02_bench_c3_w4.err
echo "bench c3 rc=$?"; tail -2 gpurun_out/r02_bench_c3_w4.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_tp2_os1.json", "gpurun_out/r02_bench_tp2_os0.json", "gpurun_out/r02_bench_c3_w4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        tp = d.get("tp_parity") or {}
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k: tp.get(k) for k in ("rel", "ok", "rel_final_logits")}, d.get("step_breakdown_us"), d["config"].get("exchange"), {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
