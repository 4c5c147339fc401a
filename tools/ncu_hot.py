"""Top stall sites of one kernel from an ncu report's source page (SASS view).
    python tools/ncu_hot.py REPORT.ncu-rep KERNEL_REGEX [N]
"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# the output may hold several kernel instances: take the first
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
end = len(rows)
for i in range(start + 1, len(rows)):
    if rows[i] and rows[i][0] == "Kernel Name":
        end = i
        break
body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
si = hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = sum(int(r[si] or 0) for r in body)
print(f"{len(body)} SASS instructions, {tot} samples")
agg = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    top = sorted(((int(r[j] or 0), hdr[j]) for j in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}%  {r[1].strip()[:90]:90s} {top}")

if len(sys.argv) > 4:   # bucketed view: samples per `bucket` consecutive instructions
    b = int(sys.argv[4])
    print(f"--- samples per {b} instructions")
    for i in range(0, len(body), b):
        chunk = body[i:i + b]
        sm = sum(int(r[si] or 0) for r in chunk)
        if sm >= tot * 0.01:
            ag = {}
            for r in chunk:
                for j in stalls:
                    if "Not Issued" not in hdr[j]:
                        ag[hdr[j]] = ag.get(hdr[j], 0) + int(r[j] or 0)
            top = sorted(ag.items(), key=lambda kv: -kv[1])[:3]
            ops = {}
            for r in chunk:
                op = r[1].strip().split()[0] if not r[1].strip().startswith("@") else r[1].strip().split()[1]
                ops[op] = ops.get(op, 0) + 1
            print(f"{i:6d} {sm:6d} {100*sm/tot:5.1f}% {top} ops={sorted(ops.items(), key=lambda kv: -kv[1])[:4]}")
