#!/bin/bash
# new standalone-stage tests first, then the whole GPU suite, then stage times of the 6M scene
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q > gpurun_out/pytest_stages.txt 2>&1
tail -30 gpurun_out/pytest_stages.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.txt 2>&1
tail -5 gpurun_out/pytest_all.txt
python scripts/stage_times.py --n 6000000 --cams outside inside 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cam'], round(d['frame_ms'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})"
