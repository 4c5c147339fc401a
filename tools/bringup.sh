#!/bin/bash
# Runs the bring-up cases one per process (a device trap must not poison the next case).
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
run() { echo "=== $*" >> $LOG; timeout 180 python tools/bringup_gemm.py "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run gemm 128 128 128 128
run gemm 128 256 512 256
run gemm 128 128 4096 128
run gemm 512 4096 4096 128
run gemm 512 4096 4096 256
run gemm 32 4096 4096 128
run gemm 200 1000 1008 128
run dequant 512 12288 4096 128
run dequant 512 12288 4096 256
run dequant 512 4096 11008 128
run fused 512 4096 4096 128 0
run fused 512 4096 4096 128 41
run fused 512 12288 4096 128 41
run fused 512 12288 4096 256 41
run fused 512 12288 4096 128 129
run fused 32 4096 4096 128 41
run fused4 512 4096 4096 128 128
run fused4 512 11008 4096 256 128
cat $LOG
