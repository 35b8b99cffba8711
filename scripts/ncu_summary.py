"""Summarise gpurun_out/*.ncu-rep + launch lists into profiles/ (tracked).  Usage:
    python scripts/ncu_summary.py <tag> [kernel ...]"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
kernels = sys.argv[2:] or ["preprocess_kernel", "onesweep4_kernel", "raster_gather4_kernel"]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


summary = {}
for k in kernels:
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{k}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    rows, units = raw_rows(rep)
    entries = []
    for d in rows:
        e = {"kernel": d.get("Kernel Name")}
        for w in WANT:
            if w in d and d[w] not in ("", "n/a"):
                e[w] = f"{d[w]} {units.get(w, '')}".strip()
        stalls = {x.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for x, v in d.items()
                  if "pcsamp_warps_issue_stalled" in x and "not_issued" not in x and v not in ("", "n/a")}
        tot = sum(stalls.values()) or 1.0
        e["stall_pct"] = {s: round(100 * v / tot, 1) for s, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]}
        entries.append(e)
    summary[k] = entries

with open(os.path.join(out_dir, f"{tag}_ncu_full_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)

# launch list: per-kernel totals and shares over the captured window
ll = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(ll):
    lines = [l for l in open(ll) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        tot[name] += float(r[vi].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    with open(os.path.join(out_dir, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): shares, not absolutes\n")
        f.write(f"# source: gpurun_out/launches_{tag}.csv ({sum(cnt.values())} launches, {total / 1e6:.3f} ms)\n")
        for name, t in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{100 * t / total:6.2f}%  {t / cnt[name] / 1e3:10.1f} us/launch  x{cnt[name]:4d}  {name}\n")
    import shutil
    shutil.copy(ll, os.path.join(out_dir, f"{tag}_launches.csv"))
print(open(os.path.join(out_dir, f"{tag}_launch_shares.txt")).read() if os.path.exists(ll) else "no launch list")
