"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).

    python tools/summarize_ncu.py TAG        # reads gpurun_out/TAG_launches.csv, gpurun_out/TAG_prof_linear.ncu-rep, TAG_bench_linear.jsonl
"""
import collections
import csv
import io
import os
import subprocess
import sys

tag = sys.argv[1]
out = [f"# ncu summary `{tag}`", ""]
lc = f"gpurun_out/{tag}_launches.csv"
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out += ["## Launch list of ONE steady-state decode step (Llama-2-7B, batch 512)",
            "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tools/profile_step.py` "
            "(per-launch times are cold-cache and serialised: compare shares, not absolutes)", "",
            "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {a[0]} | {a[1]/1e3:.1f} | {100*a[1]/tot:.1f}% | {a[1]/a[0]/1e3:.1f} |")
    out += [f"| total | {sum(a[0] for a in agg.values())} | {tot/1e3:.1f} | | |", ""]
rep = f"gpurun_out/{tag}_prof_linear.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
    names = ["W_pack", "o_proj", "swiglu_pair", "down_proj"] if len(data) == 4 else ["W_pack", "o_proj", "up_proj", "gate_proj", "down_proj"]
    out += ["## `ncu --set full` of the MixLinear kernel, layer 0 (one launch each)", "",
            "| metric | unit | " + " | ".join(names[: len(data)]) + " |", "|---|---|" + "---:|" * len(data)]
    for w in want:
        if w in idx:
            out.append(f"| `{w}` | {units[idx[w]]} | " + " | ".join(r[idx[w]] for r in data) + " |")
    out.append("")
    # DRAM bytes per launch for bench.py's roofline.traffic (average over the Linear launches of one layer)
    import json
    rd = [float(r[idx["dram__bytes_read.sum"]].replace(",", "")) for r in data]
    wr = [float(r[idx["dram__bytes_write.sum"]].replace(",", "")) for r in data]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    ru, wu = scale[units[idx["dram__bytes_read.sum"]]], scale[units[idx["dram__bytes_write.sum"]]]
    per = [a * ru + b * wu for a, b in zip(rd, wr)]
    json.dump({"dram_bytes_per_launch_avg": sum(per) / len(per), "dram_bytes_per_launch": dict(zip(names, per)),
               "source": f"profiles/{tag}_ncu_summary.md: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                         f"one launch each of {', '.join(names[:len(per)])} (Llama-2-7B layer 0, batch 512)"},
              open("profiles/ncu_traffic.json", "w"), indent=1)
bl = f"gpurun_out/{tag}_bench_linear.jsonl"
if os.path.exists(bl):
    out += ["## tools/bench_linear.py (CUDA-event times, graph of 12 launches on distinct weights, NOT under ncu)", "", "```"]
    out += [l.rstrip() for l in open(bl) if l.startswith("{")]
    out += ["```", ""]
os.makedirs("profiles", exist_ok=True)
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out))
print("\n".join(out[:40]))
