#!/bin/bash
# K1 A/B: bench scene outside / inside at 1080p, and one eighth of an 8K frame as a strip (rank 3 of 8)
for L in "$@"; do
  echo "== $L"
  export SB_LIB=$PWD/wgpu-3dgs-viewer_b200/$L/libsplat_b200.so
  python scripts/stage_times.py --n 6000000 --cams outside inside 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cam'], round(d['frame_ms'],3), 'V', d['visible'], {k:round(v,4) for k,v in d['stages_ms'].items()})"
  python scripts/stage_times.py --n 6000000 --cams outside --size 7680 4320 --strip 1616 544 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('8K strip', round(d['frame_ms'],3), 'V', d['visible'], {k:round(v,4) for k,v in d['stages_ms'].items()})"
done
