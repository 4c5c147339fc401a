#!/bin/bash
# C4 / C5 at N GPUs (after the bench fix), optionally the 7B line too (WITH7B=1)
mkdir -p gpurun_out
export MIXQ_PEER_TIMEOUT_MS=20000
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
run() { local name=$1; shift
  timeout 900 $TR --master-port 2959$N bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err
  echo "bench $name rc=$?"; python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r02_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    tp = d.get("tp_parity") or {}
    print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), {k: tp.get(k) for k in ("rel", "ok", "layers_compared")}, d.get("step_breakdown_us"), {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
except Exception as e:
    print("ERR", e)
PY
}
[ -n "$WITH7B" ] && run tp$N
run c4_tp$N --model llama-3-8b
[ "$N" = 8 ] && run c5_tp8 --model llama-2-70b --batch 128
