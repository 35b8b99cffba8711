#!/bin/bash
mkdir -p gpurun_out
for L in lib_nopipe lib; do
SB_LIB=$PWD/wgpu-3dgs-viewer_b200/$L/libsplat_b200.so python scripts/stage_times.py --n 6000000 --cams outside inside 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$L', d['cam'], round(d['frame_ms'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})"
SB_LIB=$PWD/wgpu-3dgs-viewer_b200/$L/libsplat_b200.so python scripts/stage_times.py --n 1000000 --cams outside inside 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$L 1M', d['cam'], round(d['frame_ms'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})"
done
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.txt 2>&1
tail -4 gpurun_out/pytest_all.txt
