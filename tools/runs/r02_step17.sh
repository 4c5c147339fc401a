#!/bin/bash
mkdir -p gpurun_out
BIT=4 NOUT=128 MODES=skip SHAPES=12288x4096 CADENCE=1 timeout 200 python tools/trace_linear.py > gpurun_out/r02_trace_w4b.log 2>&1
cut -c1-900 gpurun_out/r02_trace_w4b.log | tail -5
timeout 900 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x -k "w4 or int4 or quik or C3 or llama" 2>&1 | tail -3
{
timeout 300 python tools/bench_linear.py --shapes 12288x4096 --modes norm,skip --bit 4 --nout 128
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip --bit 4 --nout 128
} 2>&1 | tee gpurun_out/r02_sweep17.jsonl
