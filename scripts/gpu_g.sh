#!/bin/bash
mkdir -p gpurun_out
python scripts/config_bench.py > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_r01.err
python scripts/sort_bench.py --check --n 1000000 6000000 > gpurun_out/sort_bench_r01.jsonl 2>&1
python scripts/sort_bench.py --n 64000000 >> gpurun_out/sort_bench_r01.jsonl 2>&1
tail -3 gpurun_out/configs_r01.err; wc -l gpurun_out/configs_r01.jsonl; cat gpurun_out/sort_bench_r01.jsonl | cut -c1-150
