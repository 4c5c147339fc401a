"""Aggregate ncu stall samples per CUDA source line.
    python tools/ncu_lines.py REP [launch_skip] [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fpath = None; hdr = None; agg = collections.Counter(); text = {}; stalls = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0].isdigit():
        key = (fpath, int(r[0])); text[key] = r[1]
        try: s = int(r[hdr.index("# Samples")])
        except ValueError: s = 0
        agg[key] += s
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try: stalls[key][h] += int(r[i])
                except ValueError: pass
tot = sum(agg.values())
print("total samples", tot)
for key, s in agg.most_common(top):
    st = ", ".join(f"{k[6:]}={v}" for k, v in stalls[key].most_common(3) if v)
    print(f"{s:7d} {100*s/max(tot,1):5.1f}%  {key[0]}:{key[1]:<4d} {text[key][:90]}   [{st}]")
