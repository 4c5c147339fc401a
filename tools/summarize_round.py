"""profiles/r02_summary.md from the committed final bench lines (profiles/r02_bench_*_final.json)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*_final.json"))):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r, tp = d["roofline"], d.get("tp_parity") or {}
    pl = " / ".join(f"{k} {v['us']:.1f}" for k, v in r["per_linear"].items())
    rows.append((d["config"]["workload"].split(",")[0] + f", batch {d['config']['global_batch']}", d["n_gpus"], d["value"], d["e2e"]["value"],
                 d["ms_per_step"], f"{r['frac']:.3f} of {r['peak']:.0f} {r['unit']} ({'burst' if 'burst figure' in r['peak_source'] else 'sustained'})",
                 pl, (f"{tp.get('rel'):.1e} ok={tp.get('ok')}" if tp else "-"), os.path.basename(f)))
rows.sort(key=lambda t: (t[0], t[1]))
out = ["# Round 2 — final bench lines (bench.py on B200, one JSON line each under profiles/)", "",
       "| workload | GPUs | tokens/s | e2e tokens/s | ms/step | roofline.frac of the Linear launches | per-launch µs | tp_parity (layer-0 residual rel) | file |",
       "|---|---:|---:|---:|---:|---|---|---|---|"]
for w, n, v, e, ms, fr, pl, tp, f in rows:
    out.append(f"| {w} | {n} | {v:,.0f} | {e:,.0f} | {ms:.3f} | {fr} | {pl} | {tp} | `{f}` |")
out += ["", "Peaks: `profiles/r02_int8_peak.json` (INT8 3 040 TOPS burst / 2 607 sustained), `MEASURED_PEAKS.json` (HBM 6 426 GB/s).",
        "At N > 1 the roofline block times the four launches alone with their own prologues on the per-rank shapes; the step itself",
        "runs W_pack / the SwiGLU pair without prologue (quantised by the exchange's finish kernel) — `step_breakdown_us` in each file.", ""]
open(os.path.join(ROOT, "profiles", "r02_summary.md"), "w").write("\n".join(out))
print("\n".join(out))
