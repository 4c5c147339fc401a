#!/bin/bash
# Kernel-level numbers + ncu evidence.  TAG names the output files.
TAG=${TAG:-r01}
mkdir -p gpurun_out
( timeout 300 python tools/bench_linear.py --shapes 7b --modes norm,plain,skip;
  timeout 300 python tools/bench_linear.py --shapes 7b --nout 0 --modes plain,skip ) > gpurun_out/${TAG}_bench_linear.jsonl 2>&1
cat gpurun_out/${TAG}_bench_linear.jsonl
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mixq_linear -c 4 \
    -f -o gpurun_out/${TAG}_prof_linear python tools/profile_step.py --layers 1 > gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
NOUT=41 MODES=plain,skip SHAPES=12288x4096,4096x4096,4096x11008 timeout 200 python tools/trace_linear.py > gpurun_out/${TAG}_trace.log 2>&1
ls -la gpurun_out
