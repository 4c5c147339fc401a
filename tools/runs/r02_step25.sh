#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_k or config_shapes or w4" 2>&1 | tail -3
for cap in 1 0; do
  echo "== MIXQ_DEBUG_SPLITS=$cap (0 = planner decides)"
  MIXQ_DEBUG_SPLITS=$cap timeout 300 python tools/bench_linear.py --M 128 --shapes 1280x8192,8192x1024,3584x8192,8192x3584 --modes norm,skip --nout 41 2>&1 | grep "^{" | tee -a gpurun_out/r02_bench_linear_smallM_v3_cap$cap.jsonl | cut -c1-120
  MIXQ_DEBUG_SPLITS=$cap timeout 300 python tools/bench_linear.py --M 32 --shapes 4096x4096,4096x11008 --modes plain --nout 41 2>&1 | grep "^{" | tee -a gpurun_out/r02_bench_linear_smallM_v3_cap$cap.jsonl | cut -c1-120
done
for cap in 2; do echo "== forced $cap"; MIXQ_DEBUG_SPLITS=$cap timeout 300 python tools/bench_linear.py --M 128 --shapes 1280x8192,3584x8192,8192x3584 --modes skip --nout 41 2>&1 | grep "^{" | cut -c1-120; done
SPLITK=1 M=128 NOUT=41 MODES=skip SHAPES=1280x8192 timeout 200 python tools/trace_linear.py 2>&1 | grep -A1 "^N=" | cut -c1-600 | tee gpurun_out/r02_trace_splitk_v3.log
