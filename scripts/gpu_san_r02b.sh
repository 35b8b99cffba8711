#!/bin/bash
# compute-sanitizer over the binning kernels of the end of round 2 (dup_count / dup_offsets / dup_emit2): window-spanning splats,
# a strip render, ragged / tiny inputs
mkdir -p gpurun_out
K="default_binning or small_strict or ragged"
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_binning.py tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/san3_mem.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san3_mem.txt | tail -3
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_binning.py tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/san3_race.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/san3_race.txt | tail -3
timeout 150 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_binning.py -m gpu -q -x -k "default_binning" > gpurun_out/san3_sync.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san3_sync.txt | tail -3
