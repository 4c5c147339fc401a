#!/bin/bash
# 8 GPUs, one call: exchange correctness at 8 and 4 ranks, exchange latency anatomy at 4 / 8 ranks, then the scaling benches
# (Llama-2-7B N = 4, 8), C4 (Llama-3-8B, N = 2, 4, 8) and C5 (Llama-2-70B batch 128, N = 8).
mkdir -p gpurun_out
export MIXQ_PEER_TIMEOUT_MS=20000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CHECK_KIND=push timeout 600 $TR --nproc-per-node ${NGPU:-8} --master-port 29561 tools/check_peer_exchange.py > gpurun_out/r02_check_exchange_${NGPU:-8}.log 2>&1
echo "check rc=$?"; grep -E " ok on |Error|error" gpurun_out/r02_check_exchange_${NGPU:-8}.log | tail -8
for n in ${ANATOMY:-8}; do
  [ "$n" = none ] && continue
  timeout 400 $TR --nproc-per-node $n --master-port 2957$n tools/bench_exchange.py > gpurun_out/r02_bench_exchange_$n.json 2> gpurun_out/r02_bench_exchange_$n.err
  echo "anatomy $n rc=$?"; grep "^{" gpurun_out/r02_bench_exchange_$n.json
done
run() {  # name, nproc, extra bench args, env
  local name=$1 n=$2; shift 2
  timeout 900 $TR --nproc-per-node $n --master-port 2958$n bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err
  echo "bench $name rc=$?"; tail -1 gpurun_out/r02_bench_$name.err | cut -c1-300
}
N=${NGPU:-8}
run tp$N $N
run c4_tp$N $N --model llama-3-8b
if [ "$N" = "8" ]; then run c5_tp8 8 --model llama-2-70b --batch 128; fi
if [ "$N" = "4" ]; then MIXQ_TP_ONE_SHOT=1 run tp4_oneshot 4; fi
if [ -n "$BCAST_PEER_TOO" ]; then MIXQ_TP_BCAST=peer run tp${N}_peerbcast $N; fi
if [ -n "$FLAGS_TOO" ]; then MIXQ_TP_SYNC=flags run tp${N}_flags $N; fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_tp[248]*.json") + glob.glob("gpurun_out/r02_bench_c[45]_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        tp = d.get("tp_parity") or {}
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k: tp.get(k) for k in ("rel", "ok", "rel_final_logits", "layers_compared")}, d.get("step_breakdown_us"), d["config"].get("exchange"), {k: round(v["us"], 1) for k, v in d["roofline"]["per_linear"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
