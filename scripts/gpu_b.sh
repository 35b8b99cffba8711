#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "sorter or ragged or ties or small_strict or full_size or overbright" > gpurun_out/pytest_b.txt 2>&1
tail -5 gpurun_out/pytest_b.txt
python scripts/sort_bench.py --check --n 1000000 6000000 > gpurun_out/sort_v2.txt 2>&1
python scripts/sort_bench.py --n 64000000 >> gpurun_out/sort_v2.txt 2>&1
SB_SORT_IMPL=v1 python scripts/sort_bench.py --n 6000000 64000000 > gpurun_out/sort_v1.txt 2>&1
echo v2; cat gpurun_out/sort_v2.txt; echo v1; cat gpurun_out/sort_v1.txt
python scripts/stage_times.py --n 6000000 --cams outside > gpurun_out/stage_v2.txt 2>&1; cat gpurun_out/stage_v2.txt
