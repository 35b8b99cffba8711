#!/bin/bash
bash scripts/profile.sh r02 > gpurun_out/profile_r02.log 2>&1
tail -c 400 gpurun_out/bench_r02.json
python scripts/strips_8k.py > gpurun_out/strips_n1.txt 2>&1; tail -1 gpurun_out/strips_n1.txt | cut -c1-250
python scripts/sort_bench.py > gpurun_out/sort_bench_r02.jsonl 2>/dev/null; cat gpurun_out/sort_bench_r02.jsonl | cut -c1-200
