#!/bin/bash
mkdir -p gpurun_out
BIT=4 NOUT=128 MODES=skip SHAPES=12288x4096 CADENCE=1 timeout 200 python tools/trace_linear.py > gpurun_out/r02_trace_w4.log 2>&1
cut -c1-1500 gpurun_out/r02_trace_w4.log
