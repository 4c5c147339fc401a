#!/bin/bash
mkdir -p gpurun_out
for c in ab8e1e4 v1 v2_now4 v3_nobcast v4_both ab8e1e4 v1; do
  L=gpurun_ab/libmixq_$c.so
  echo "== $c"
  MIXQ_LIB=$L MIXQ_LIB_LENIENT=1 timeout 300 python tools/bench_linear.py --shapes 12288x4096,4096x4096 --modes norm,skip 2>&1 | grep -o '"N.*'
done 2>&1 | tee gpurun_out/r02_ab12.log
