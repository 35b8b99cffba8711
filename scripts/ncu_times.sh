#!/bin/bash
# Per-kernel durations (ncu, cold-cache / serialised: compare, do not quote) of one camera of the 6 M bench scene.
# Usage: ncu_times.sh <kernel-regex> <cam> ["ENV=.."] [extra stage_times args]
RX=$1; CAM=$2; E=$3; shift 3
mkdir -p gpurun_out
env $E ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$RX" --csv --log-file gpurun_out/ncu_times.csv \
    python scripts/stage_times.py --n 6000000 --cams $CAM --iters 3 "$@" > /dev/null 2>&1
python - <<'P'
import csv, collections, statistics
rows = [r for r in csv.reader(open("gpurun_out/ncu_times.csv")) if len(r) > 10]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
t = collections.OrderedDict()
for r in rows[1:]:
    t.setdefault(r[k][:60], []).append(float(r[v].replace(",", "")))
for name, xs in t.items():
    print(f"  {statistics.median(xs)/1000:9.1f} us  x{len(xs):3d}  {name}")
P
