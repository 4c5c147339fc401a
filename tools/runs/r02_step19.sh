#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x 2>&1 | tail -3
{
timeout 300 python tools/bench_linear.py --shapes 12288x4096,4096x4096,4096x11008 --modes norm,plain,skip
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip
timeout 300 python tools/bench_linear.py --shapes 12288x4096 --modes norm,skip --bit 4 --nout 128
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip --bit 4 --nout 128
timeout 100 python tools/bench_attn.py
} 2>&1 | tee gpurun_out/r02_sweep19.jsonl
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --kv-len 0 > gpurun_out/r02_bench_step19.json 2> gpurun_out/r02_bench_step19.err
timeout 600 python bench.py --bit 4 --steps 20 --warmup 5 --no-cpu-baseline --kv-len 0 > gpurun_out/r02_bench_c3_w4_v5.json 2> gpurun_out/r02_bench_c3_w4_v5.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_step19.json", "gpurun_out/r02_bench_c3_w4_v5.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
