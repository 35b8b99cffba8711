#!/bin/bash
# Usage: scripts/profile.sh <round-tag>   (runs under gpurun; writes gpurun_out/)
TAG=${1:-r02}
mkdir -p gpurun_out
set -x
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json
# launch list (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
# full captures of the three hot kernels
for K in preprocess_kernel onesweep4_kernel raster_gather4_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/prof_${K}_$TAG \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out
