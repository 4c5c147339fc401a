#!/bin/bash
# quick correctness + kernel timing loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/bench_linear.py --shapes ${SHAPES:-7b} ${BL_ARGS} 2>&1 | tee gpurun_out/bench_linear.jsonl
