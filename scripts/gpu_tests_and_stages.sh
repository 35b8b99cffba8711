#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_e.txt 2>&1
tail -4 gpurun_out/pytest_e.txt
python scripts/stage_times.py --n 6000000 --cams outside inside --count 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cam'], round(d['frame_ms'],3), {k:round(v,3) for k,v in d['stages_ms'].items()}, d.get('counters'))"
