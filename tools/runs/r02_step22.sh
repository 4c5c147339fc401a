#!/bin/bash
# split-K: parity tests, then the small-M shapes with and without the workspace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for flag in "--no-splitk" ""; do
  echo "== C5 per-rank shapes (70B TP8, M=128) $flag"
  timeout 300 python tools/bench_linear.py --M 128 --shapes 1280x8192,8192x1024,3584x8192,8192x3584 --modes norm,plain,skip --nout 41 $flag 2>&1 | grep "^{" | tee -a gpurun_out/r02_bench_linear_smallM${flag:+_nosplit}.jsonl | cut -c1-200
  echo "== C1 (M=32, 4096x4096) $flag"
  timeout 300 python tools/bench_linear.py --M 32 --shapes 4096x4096,12288x4096,11008x4096,4096x11008 --modes plain,skip --nout 41 $flag 2>&1 | grep "^{" | tee -a gpurun_out/r02_bench_linear_smallM${flag:+_nosplit}.jsonl | cut -c1-200
done
