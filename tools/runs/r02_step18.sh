#!/bin/bash
mkdir -p gpurun_out
BIT=4 NOUT=128 MODES=skip SHAPES=12288x4096 CADENCE=1 timeout 200 python tools/trace_linear.py > gpurun_out/r02_trace_w4c.log 2>&1
grep "mma" gpurun_out/r02_trace_w4c.log | cut -c1-400
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x 2>&1 | tail -3
{
timeout 300 python tools/bench_linear.py --shapes 12288x4096 --modes norm,skip --bit 4 --nout 128
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip --bit 4 --nout 128
} 2>&1 | tee gpurun_out/r02_sweep18.jsonl
timeout 600 python bench.py --bit 4 --steps 20 --warmup 5 --no-cpu-baseline --kv-len 0 > gpurun_out/r02_bench_c3_w4_v4.json 2> gpurun_out/r02_bench_c3_w4_v4.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_c3_w4_v4.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
