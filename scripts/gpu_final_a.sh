#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.txt 2>&1
tail -6 gpurun_out/pytest_all.txt
python scripts/config_bench.py > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_err.txt
tail -3 gpurun_out/configs_err.txt
wc -l gpurun_out/configs_r01.jsonl
