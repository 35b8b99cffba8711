#!/bin/bash
# A few counters of one kernel, for each library variant.  Usage: ncu_metrics.sh <kernel-regex> <cam> <libdir> ...
RX=$1; CAM=$2; shift 2
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__thread_inst_executed.sum,launch__registers_per_thread,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum
mkdir -p gpurun_out
for L in "$@"; do
  echo "== $L"
  SB_LIB=$PWD/wgpu-3dgs-viewer_b200/$L/libsplat_b200.so ncu --metrics $M --clock-control none -k "regex:$RX" -s 3 -c 1 --csv --log-file gpurun_out/ncu_m.csv \
      python scripts/stage_times.py --n 6000000 --cams $CAM --iters 2 > /dev/null 2>&1
  python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/ncu_m.csv")) if len(r) > 10]
h = rows[0]; n = h.index("Metric Name"); v = h.index("Metric Value")
for r in rows[1:]:
    print(f"   {r[n]:75s} {r[v]}")
P
done
