#!/bin/bash
# A/B of library build variants: per-stage times on the bench scene.  Usage: gpu_variants.sh <libdir> ...
mkdir -p gpurun_out
for L in "$@"; do
  echo "== $L"
  SB_LIB=$PWD/wgpu-3dgs-viewer_b200/$L/libsplat_b200.so python scripts/stage_times.py --n 6000000 --cams outside inside 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cam'], round(d['frame_ms'],3), {k:round(v,4) for k,v in d['stages_ms'].items()})"
done
