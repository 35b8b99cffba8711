#!/bin/bash
echo votes; python scripts/sort_bench.py --n 6000000 64000000 --iters 10 2>&1 | grep u32 | cut -c1-120
echo match; SB_LIB=$PWD/wgpu-3dgs-viewer_b200/lib_m/libsplat_b200.so python scripts/sort_bench.py --check --n 6000000 2>&1 | grep u32 | cut -c1-140
SB_LIB=$PWD/wgpu-3dgs-viewer_b200/lib_m/libsplat_b200.so python scripts/sort_bench.py --n 64000000 --iters 10 2>&1 | grep u32 | cut -c1-120
