#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for t in 0 96 128 160 192 224 256; do python tools/bench_linear.py --shapes ${SHAPES:-7b} --modes ${MODES:-skip,plain} --tile $t; done 2>&1 | tee gpurun_out/sweep.jsonl
