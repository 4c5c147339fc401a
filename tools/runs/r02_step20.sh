#!/bin/bash
# Round-2 evidence on one GPU: GPU tests, the bench line (W8 headline, W4 C3, KV variant is inside), ncu launch list + full captures.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_n1_final.json
timeout 600 python bench.py --bit 4 > gpurun_out/r02_bench_c3_w4_final.json 2> gpurun_out/r02_bench_c3_w4_final.err; echo "bench w4 rc=$?"; cut -c1-300 gpurun_out/r02_bench_c3_w4_final.json
TAG=r02 bash tools/gpu_profile.sh 2>&1 | tail -30
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mixq_linear -c 4 \
    -f -o gpurun_out/r02w4_prof_linear python tools/profile_step.py --layers 1 --bit 4 > gpurun_out/r02w4_prof.log 2>&1
tail -2 gpurun_out/r02w4_prof.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rope_attn -c 1 \
    -f -o gpurun_out/r02_prof_attn python tools/profile_step.py --layers 1 > gpurun_out/r02_prof_attn.log 2>&1
tail -2 gpurun_out/r02_prof_attn.log
timeout 300 python tools/bench_linear.py --shapes 7b --modes norm,skip --bit 4 --nout 128 > gpurun_out/r02_bench_linear_w4.jsonl 2>&1; cat gpurun_out/r02_bench_linear_w4.jsonl
ls -la gpurun_out | head -50
