#!/bin/bash
# Usage: scripts/ncu_frame.sh <tag> <kernel-regex> [count] [extra stage_times args] -> gpurun_out/prof_sel_<tag>.ncu-rep (--set full)
TAG=$1; RX=$2; CNT=${3:-8}; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $CNT -c $CNT -f -o gpurun_out/prof_sel_$TAG \
    python scripts/stage_times.py --n 6000000 --cams outside --iters 2 "$@" > gpurun_out/ncu_sel_$TAG.log 2>&1
tail -2 gpurun_out/ncu_sel_$TAG.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
