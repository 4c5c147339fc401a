#!/bin/bash
# round 2, step 0: INT8 peak, MMA issue rates, baseline GPU tests, tile sweeps of the existing kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r02_nvsmi.txt
timeout 600 python tools/int8_peak.py --out gpurun_out/r02_int8_peak.json > gpurun_out/r02_int8_peak.log 2>&1
timeout 120 ./tools/mma_bw > gpurun_out/r02_mma_bw.log 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
{
for t in 0 192 224 256; do timeout 300 python tools/bench_linear.py --shapes 12288x4096 --modes norm,skip --tile $t; done
for t in 0 192 224 256; do timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip --tile $t; done
timeout 300 python tools/bench_linear.py --shapes 4096x4096,4096x11008 --modes plain,skip
timeout 300 python tools/bench_linear.py --shapes 12288x4096,4096x4096,4096x11008 --modes skip --nout 0
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairskip --nout 0
} 2>&1 | tee gpurun_out/r02_sweep0.jsonl
