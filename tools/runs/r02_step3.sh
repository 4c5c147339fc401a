#!/bin/bash
# attention kernels alone + ncu source-level captures of the attention+quant kernel and of the SwiGLU pair / W_pack GEMM
mkdir -p gpurun_out
timeout 200 python tools/bench_attn.py 2>&1 | tee gpurun_out/r02_bench_attn.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rope_attn -s 4 -c 2 -f -o gpurun_out/r02_prof_attn \
    python tools/bench_attn.py --eager --copies 2 > gpurun_out/r02_prof_attn.log 2>&1
tail -2 gpurun_out/r02_prof_attn.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mixq_linear2 -s 2 -c 1 -f -o gpurun_out/r02_prof_pair \
    python tools/bench_linear.py --eager --copies 2 --shapes 11008x4096 --modes pairskip > gpurun_out/r02_prof_pair.log 2>&1
tail -2 gpurun_out/r02_prof_pair.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mixq_linear2 -s 2 -c 1 -f -o gpurun_out/r02_prof_wpack \
    python tools/bench_linear.py --eager --copies 2 --shapes 12288x4096 --modes skip > gpurun_out/r02_prof_wpack.log 2>&1
tail -2 gpurun_out/r02_prof_wpack.log
ls -la gpurun_out/*.ncu-rep
