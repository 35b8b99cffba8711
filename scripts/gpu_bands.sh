#!/bin/bash
mkdir -p gpurun_out
for B in 1 2 4 8 17; do
  SB_RASTER_BANDS=$B python bench.py --steps 32 --warmup 6 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('bands', $B, 'batch', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', round(d['single_frame']['frames_per_s'],1), 'raster_ms', round(d['stages']['ms']['raster'],3))"
done
