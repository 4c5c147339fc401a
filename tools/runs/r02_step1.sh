#!/bin/bash
# round 2, step 1: attention-fused o_proj prologue, new config-shape tests, pair tile width 160, no-outlier upper bounds
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
{
for t in 160 128; do timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip --tile $t; done
for t in 160 224 0; do timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip --tile $t --nout 0; done
for t in 192 0; do timeout 300 python tools/bench_linear.py --shapes 12288x4096 --modes norm,skip --tile $t --nout 0; done
} 2>&1 | tee gpurun_out/r02_sweep1.jsonl
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_attnquant.json 2> gpurun_out/r02_bench_attnquant.err
tail -c 1500 gpurun_out/r02_bench_attnquant.json
MIXQ_FUSE_ATTN_QUANT=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_noattnquant.json 2>> gpurun_out/r02_bench_attnquant.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_attnquant.json", "gpurun_out/r02_bench_noattnquant.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
