#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_step2.json 2> gpurun_out/r02_bench_step2.err
MIXQ_FUSE_ATTN_QUANT=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_step2_noaq.json 2>> gpurun_out/r02_bench_step2.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_step2.json", "gpurun_out/r02_bench_step2_noaq.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["value_median"]), round(d["e2e"]["value"]), d["roofline"]["frac"], d["roofline"]["peak_source"][:60], {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r02_bench_step2.err
