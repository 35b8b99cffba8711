#!/bin/bash
# one GPU session: probe, parity tests, A/B stage times, fresh ncu captures
mkdir -p gpurun_out
scripts/probes/probe_ffma2 > gpurun_out/ffma2.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.txt 2>&1
tail -5 gpurun_out/pytest_a.txt
python scripts/stage_times.py --n 6000000 --cams outside inside --count > gpurun_out/stage_obb.txt 2>&1
SB_RASTER_CULL=bbox python scripts/stage_times.py --n 6000000 --cams outside inside --count > gpurun_out/stage_bbox.txt 2>&1
cat gpurun_out/ffma2.txt gpurun_out/stage_obb.txt gpurun_out/stage_bbox.txt
scripts/ncu_sort.sh r01b 6000000
scripts/ncu_one.sh raster_gather4_kernel r01b
