#!/bin/bash
# usage: ncu_two.sh <kernel-regex> ...   full ncu capture of one launch of each kernel -> gpurun_out/prof_<kernel>_x.ncu-rep
mkdir -p gpurun_out
for K in "$@"; do
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_x \
    python scripts/stage_times.py --n 6000000 --cams outside --iters 2 > gpurun_out/ncu_${K}_x.log 2>&1
done
