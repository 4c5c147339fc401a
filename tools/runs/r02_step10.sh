#!/bin/bash
mkdir -p gpurun_out
{
echo "== W_pack"; NOUT=41 MODES=skip,plain SHAPES=12288x4096 CADENCE=1 timeout 200 python tools/trace_linear.py
echo "== pair nout 41"; PAIR=1 NOUT=41 MODES=skip SHAPES=11008x4096 CADENCE=1 timeout 200 python tools/trace_linear.py
echo "== pair nout 0"; PAIR=1 NOUT=0 MODES=skip SHAPES=11008x4096 CADENCE=1 timeout 200 python tools/trace_linear.py
echo "== o_proj / down"; NOUT=41 MODES=skip SHAPES=4096x4096,4096x11008 CADENCE=1 timeout 200 python tools/trace_linear.py
} > gpurun_out/r02_trace10.log 2>&1
timeout 300 python tools/bench_linear.py --shapes 12288x4096,4096x4096,4096x11008 --modes norm,skip 2>&1 | tee gpurun_out/r02_sweep10.jsonl
timeout 300 python tools/bench_linear.py --shapes 11008x4096 --modes pairnorm,pairskip 2>&1 | tee -a gpurun_out/r02_sweep10.jsonl
