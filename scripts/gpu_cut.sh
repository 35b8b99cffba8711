#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cutoff or counters or bbox_only" > gpurun_out/pytest_cut.txt 2>&1
tail -30 gpurun_out/pytest_cut.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.txt 2>&1
tail -5 gpurun_out/pytest_all.txt
python scripts/stage_times.py --n 6000000 --cams outside inside --count 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cam'], round(d['frame_ms'],3), {k:round(v,3) for k,v in d['stages_ms'].items()}, d.get('counters'), d.get('stats'))"
