#!/bin/bash
mkdir -p gpurun_out
export MIXQ_PEER_TIMEOUT_MS=20000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/bench_exchange.py > gpurun_out/r02_bench_exchange_2.json 2> gpurun_out/r02_bench_exchange_2.err
echo "rc=$?"; cat gpurun_out/r02_bench_exchange_2.json; tail -5 gpurun_out/r02_bench_exchange_2.err
for ab in 1 2 6; do
MIXQ_DEBUG_XF=$ab timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/bench_exchange.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ablate $ab', {k: v for k, v in d.items() if 'push' in k})
"
done
