#!/bin/bash
# final single-GPU verification: GPU tests, smoke(), the bench line (both arms)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_n1_final2.json 2> gpurun_out/r02_bench_n1_final2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_final2.json").read().strip().splitlines()[-1])
r = d["roofline"]
print(round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"], 3), "of", round(r["peak"]), {k: round(v["us"], 1) for k, v in r["per_linear"].items()}, d["clocks"], "launches", d["gpu_launches"], "traffic", r["traffic"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
