#!/bin/bash
# Usage: scripts/ncu_frame.sh <tag> [extra stage_times args] -> gpurun_out/prof_frame_<tag>.ncu-rep (--set full, ~2 whole frames of kernels)
TAG=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -s 60 -c 40 -f -o gpurun_out/prof_frame_$TAG \
    python scripts/stage_times.py --n 6000000 --cams outside --iters 4 "$@" > gpurun_out/ncu_frame_$TAG.log 2>&1
tail -3 gpurun_out/ncu_frame_$TAG.log
ncu -i gpurun_out/prof_frame_$TAG.ncu-rep --page raw --csv --metrics gpu__time_duration.sum 2>/dev/null | cut -d, -f5,12- | head -50
