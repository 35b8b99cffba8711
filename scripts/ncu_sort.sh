#!/bin/bash
# Usage: scripts/ncu_sort.sh <tag> [n] [kernel-regex ...]
TAG=$1; N=${2:-16000000}; shift 2
KS=${@:-onesweep2_kernel histogram_kernel}
mkdir -p gpurun_out
for K in $KS; do
ncu --set full --clock-control none --import-source on -k regex:$K -s 9 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
    python scripts/sort_bench.py --n $N --iters 1 > gpurun_out/ncu_${K}_$TAG.log 2>&1
tail -2 gpurun_out/ncu_${K}_$TAG.log
done
