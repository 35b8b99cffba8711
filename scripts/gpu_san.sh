#!/bin/bash
mkdir -p gpurun_out
K="small_strict or ragged or huge or selection_mask or multi_model or standalone or rect or brush or depth_stencil or overbright or ties"
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/san_mem.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_mem.txt | tail -3
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small_strict or ragged or depth_stencil" > gpurun_out/san_sync.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_sync.txt | tail -3
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged or standalone" > gpurun_out/san_race.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/san_race.txt | tail -3
timeout 600 compute-sanitizer --tool memcheck python scripts/stage_times.py --n 6000000 --cams outside --iters 1 > gpurun_out/san_6m.txt 2>&1
grep -E "ERROR SUMMARY" gpurun_out/san_6m.txt | tail -1
