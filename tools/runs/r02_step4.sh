#!/bin/bash
# 2 GPUs: attention microbench (new kernel), full gpu tests incl. the 2-rank exchange check (peer + NVLS multicast), TP=2 bench
mkdir -p gpurun_out
timeout 200 python tools/bench_attn.py 2>&1 | tee gpurun_out/r02_bench_attn2.jsonl
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
for ex in auto peer; do
MIXQ_TP_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_tp2_$ex.json 2> gpurun_out/r02_bench_tp2_$ex.err
echo "bench tp2 $ex rc=$?"; tail -3 gpurun_out/r02_bench_tp2_$ex.err
done
MIXQ_FUSE_ATTN_QUANT=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_step4_aq.json 2> gpurun_out/r02_bench_step4.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_tp2_auto.json", "gpurun_out/r02_bench_tp2_peer.json", "gpurun_out/r02_bench_step4_aq.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("tp_parity"), d.get("step_breakdown_us"), d["config"].get("exchange"))
    except Exception as e:
        print(f, "ERR", e)
PY
