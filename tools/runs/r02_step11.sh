#!/bin/bash
mkdir -p gpurun_out
for c in b21b99b 8317277 ab8e1e4 current b21b99b current; do
  if [ "$c" = "current" ]; then L=mixq_b200/lib/libmixq_sm100.so; else L=gpurun_ab/libmixq_$c.so; fi
  echo "== $c"
  MIXQ_LIB=$L MIXQ_LIB_LENIENT=1 timeout 300 python tools/bench_linear.py --shapes 12288x4096,4096x4096 --modes norm,skip 2>&1 | grep -o '"N.*'
  MIXQ_LIB=$L MIXQ_LIB_LENIENT=1 timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairskip 2>&1 | grep -o '"N.*'
done 2>&1 | tee gpurun_out/r02_ab11.log
