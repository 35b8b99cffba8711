"""Aggregate warp-stall samples of an ncu report per CUDA source line.  Usage: ncu_lines.py <rep> [top]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None
agg = defaultdict(lambda: [0.0, 0.0, "", defaultdict(float)])
hdr = None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Name":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        wi, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        w = float(r[wi])
    except ValueError:
        continue
    key = (cur_file, r[0])
    a = agg[key]
    if r[2] == "" or a[2] == "":  # the CUDA line row carries the source text
        a[2] = a[2] or r[1]
    if r[2] != "":  # sass row
        a[0] += w
        try:
            a[1] += float(r[ei])
        except ValueError:
            pass
        for i, name in stall_cols:
            try:
                a[3][name] += float(r[i])
            except ValueError:
                pass
tot = sum(a[0] for a in agg.values()) or 1
print(f"total samples {tot:.0f}")
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = sorted(a[3].items(), key=lambda x: -x[1])[:3]
    sts = " ".join(f"{k}={100 * v / max(a[0], 1):.0f}%" for k, v in st if v > 0)
    print(f"{100 * a[0] / tot:5.1f}%  {f}:{ln:>4s} inst={a[1]:>10.0f}  [{sts}]  {a[2].strip()[:90]}")
