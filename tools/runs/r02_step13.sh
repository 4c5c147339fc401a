#!/bin/bash
mkdir -p gpurun_out
for c in ab8e1e4 v5_tmpl ab8e1e4 v5_tmpl; do
  L=gpurun_ab/libmixq_$c.so
  echo "== $c"
  MIXQ_LIB=$L MIXQ_LIB_LENIENT=1 timeout 300 python tools/bench_linear.py --shapes 12288x4096,4096x4096,4096x11008 --modes norm,skip 2>&1 | grep -o '"N.*'
  MIXQ_LIB=$L MIXQ_LIB_LENIENT=1 timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip 2>&1 | grep -o '"N.*'
  MIXQ_LIB=$L MIXQ_LIB_LENIENT=1 timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairskip --nout 0 2>&1 | grep -o '"N.*'
done 2>&1 | tee gpurun_out/r02_ab13.log
timeout 300 python tools/bench_linear.py --shapes 12288x4096 --modes norm,skip --bit 4 --nout 128 2>&1 | tee -a gpurun_out/r02_ab13.log
timeout 900 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x 2>&1 | tail -3
