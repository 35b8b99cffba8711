#!/bin/bash
# sanitizer pass over the kernels changed late in round 1 (CTA-level cull barrier, 64-byte gather rows, tile schedule, vertex stage)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -m gpu -q -x -k "small_strict or ragged or cutoff or standalone or callers_indices or chain or depth_stencil or multi_model" > gpurun_out/san2_mem.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san2_mem.txt | tail -3
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -m gpu -q -x -k "small_strict or ragged or cutoff or callers_indices or depth_stencil" > gpurun_out/san2_race.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/san2_race.txt | tail -3
grep -E "Warning: Race|Error: Race" gpurun_out/san2_race.txt | sed 's/0x[0-9a-f]*//g' | cut -c1-200 | sort | uniq -c | sort -rn | head -5
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small_strict or ragged or depth_stencil" > gpurun_out/san2_sync.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san2_sync.txt | tail -3
