#!/bin/bash
# Usage: scripts/profile.sh <round-tag>   (runs under gpurun; writes gpurun_out/)
TAG=${1:-r02}
mkdir -p gpurun_out
set -x
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 1500 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 600 gpurun_out/bench_ref_$TAG.json
# launch list (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strips > gpurun_out/ncu_bench_$TAG.log 2>&1
# full captures of the three hot kernels (no source import: the reports must stay under the 64 MiB that travels back)
for K in preprocess_kernel onesweep4_kernel raster_gather4_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/prof_${K}_$TAG \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strips > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
# the binning kernels of one frame (dup_count, dup_offsets, dup_emit2)
ncu --set full --clock-control none --import-source on -k regex:dup_ -s 6 -c 3 -f -o gpurun_out/prof_binning_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strips > gpurun_out/ncu_binning_$TAG.log 2>&1
python scripts/config_bench.py > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; tail -c 600 gpurun_out/configs_$TAG.jsonl
python scripts/sort_bench.py > gpurun_out/sort_bench_$TAG.jsonl 2>/dev/null; cut -c1-160 gpurun_out/sort_bench_$TAG.jsonl
ls -la gpurun_out
