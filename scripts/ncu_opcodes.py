"""Executed-instruction histogram by SASS opcode of an ncu report (first kernel).  Usage: ncu_opcodes.py <rep> [top]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr = None
ops = defaultdict(float)
tot = 0.0
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if "Source" in r and "Instructions Executed" in r:
        hdr = r
        si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        e = float(r[ei])
    except ValueError:
        continue
    toks = r[si].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";")
    ops[op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMS", "BAR") else "")] += e
    tot += e
print(f"total warp-level instructions executed: {tot:.0f}")
for op, e in sorted(ops.items(), key=lambda x: -x[1])[:top]:
    print(f"{100 * e / tot:6.2f}%  {e:14.0f}  {op}")
