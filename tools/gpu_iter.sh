#!/bin/bash
# one development iteration: GPU parity tests, kernel-level times, optional step bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_linear.py --shapes 7b --modes norm,plain,skip 2>&1 | tee gpurun_out/iter_bench_linear.log
timeout 300 python tools/bench_linear.py --shapes 7b --modes skip --nout 0 2>&1 | tee -a gpurun_out/iter_bench_linear.log
if [ -n "$BENCH" ]; then timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_iter.err | tee gpurun_out/bench_iter.log | cut -c1-300; tail -2 gpurun_out/bench_iter.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_iter.log').read().strip().splitlines()[-1])
print({k:round(v['us'],1) for k,v in d['roofline']['per_linear'].items()}, d['roofline']['frac'])
PY
fi
