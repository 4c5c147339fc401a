#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
{
timeout 300 python tools/bench_linear.py --shapes 12288x4096,4096x4096,4096x11008 --modes norm,skip
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip --nout 0
timeout 300 python tools/bench_linear.py --shapes 12288x4096 --modes norm,skip --bit 4 --nout 128
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip --bit 4 --nout 128
} 2>&1 | tee gpurun_out/r02_sweep8.jsonl
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_step8.json 2> gpurun_out/r02_bench_step8.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_step8.err
timeout 600 python bench.py --bit 4 --steps 20 --warmup 5 --no-cpu-baseline --kv-len 0 > gpurun_out/r02_bench_c3_w4_v2.json 2> gpurun_out/r02_bench_c3_w4_v2.err
echo "bench c3 rc=$?"; tail -2 gpurun_out/r02_bench_c3_w4_v2.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_step8.json", "gpurun_out/r02_bench_c3_w4_v2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), d.get("kv_variant"), d.get("cpu_baseline"), {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
