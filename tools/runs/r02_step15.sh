#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/bench_linear.py --shapes 512x4096 --modes skip --bit 4 --nout 128 --copies 1 --eager > gpurun_out/r02_sanitizer_w4.log 2>&1
grep -E "Invalid|Illegal|error|Error|at .*cu|=========     at|by thread|Address" gpurun_out/r02_sanitizer_w4.log | head -40
