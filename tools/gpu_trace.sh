#!/bin/bash
# per-phase timelines (in-kernel %globaltimer) of the MixLinear kernel for the 7B shapes
mkdir -p gpurun_out
SHAPES=${SHAPES:-12288x4096,4096x4096,11008x4096,4096x11008} python tools/trace_linear.py 2>&1 | tee gpurun_out/${TAG:-r01}_trace.log
