#!/bin/bash
# PDL on/off: parity tests, kernel-level times, epilogue clock profile (library built with EXTRA=-DMIXQ_EPI_PROFILE)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for pdl in 1 0; do echo "MIXQ_PDL=$pdl"; MIXQ_PDL=$pdl timeout 300 python tools/bench_linear.py --shapes 7b --modes norm,skip; done 2>&1 | tee gpurun_out/pdl_bench_linear.log
for nout in 0 41; do echo "NOUT=$nout"; NOUT=$nout MODES=skip SHAPES=12288x4096,4096x4096,4096x11008 timeout 300 python tools/trace_linear.py; done 2>&1 | tee gpurun_out/epi_profile.log
for pdl in 1 0; do MIXQ_PDL=$pdl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_pdl$pdl.err | tee gpurun_out/bench_pdl$pdl.log | cut -c1-400; tail -2 gpurun_out/bench_pdl$pdl.err; done
