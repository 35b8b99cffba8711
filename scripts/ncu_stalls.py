"""Print the key metrics + stall reasons + hottest source lines of an ncu report."""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
d = dict(zip(rows[0], rows[2]))
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg"]:
    print(f"{k:70s} {d.get(k)}")
st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for k, v in d.items()
      if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and v not in ("", "n/a")}
tot = sum(st.values()) or 1
print("stalls:", {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]})
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    si, wi = hdr.index("Source"), hdr.index("# Samples")
    ls = hdr.index("stall_long_sb")
    items = []
    for r in rows[2:]:
        try:
            items.append((float(r[wi]), r[si], r[hdr.index("Instructions Executed")], r[ls]))
        except (ValueError, IndexError):
            pass
    tot = sum(x[0] for x in items) or 1
    for w, s_, ex, lsb in sorted(items, key=lambda x: -x[0])[:int(sys.argv[2])]:
        print(f"{100 * w / tot:5.1f}%  exec={ex:>9s} long_sb={lsb:>6s}  {s_[:120]}")
