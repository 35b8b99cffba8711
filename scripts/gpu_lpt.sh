#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.txt 2>&1
tail -4 gpurun_out/pytest_all.txt
for O in rows lpt; do
SB_RASTER_ORDER=$O python scripts/stage_times.py --n 6000000 --cams outside inside 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$O', d['cam'], round(d['frame_ms'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})"
done
for B in 1 2; do
  SB_RASTER_BANDS=$B python bench.py --steps 32 --warmup 6 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('bands', $B, 'batch', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', round(d['single_frame']['frames_per_s'],1), 'raster_ms', round(d['stages']['ms']['raster'],3))"
done
