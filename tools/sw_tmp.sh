for sb in 24576 32768 49152; do echo "stage_bytes $sb"; MIXQ_DEBUG_STAGE_BYTES=$sb python tools/bench_linear.py --shapes 4096x4096 --modes skip --tile 128 --nout 0; done
