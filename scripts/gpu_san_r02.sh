#!/bin/bash
# compute-sanitizer over the round-2 kernels: key-adaptive depth sort (every frame), strip culling K1 instantiation, strip-relative
# tile ids, sRGB rasterizer, colour modifiers, tile-row work, and (2 GPUs) the peer-mapped frame + partitioned strips
mkdir -p gpurun_out
K="small_strict or ragged or strip_cull or modes_and_targets or srgb or ties or selection_mask or multi_model"
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_reference_e2e.py -m gpu -q -x -k "$K or modifiers or override" > gpurun_out/san2_mem.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san2_mem.txt | tail -3
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small_strict or ragged or strip_cull or srgb" > gpurun_out/san2_race.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/san2_race.txt | tail -3
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small_strict or strip_cull or srgb" > gpurun_out/san2_sync.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san2_sync.txt | tail -3
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 900 compute-sanitizer --tool memcheck --target-processes all python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "strips_over_two" > gpurun_out/san2_mem_2gpu.txt 2>&1
  grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san2_mem_2gpu.txt | tail -4
fi
