#!/bin/bash
# Usage: scripts/ncu_one.sh <kernel-regex> <tag> [extra stage_times args]  -> gpurun_out/prof_<kernel>_<tag>.ncu-rep
K=$1; TAG=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
    python scripts/stage_times.py --n 6000000 --cams outside --iters 2 "$@" > gpurun_out/ncu_${K}_$TAG.log 2>&1
tail -3 gpurun_out/ncu_${K}_$TAG.log
