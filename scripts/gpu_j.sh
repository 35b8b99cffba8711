#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "depth_stencil" > gpurun_out/san_race2.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/san_race2.txt | tail -3
