#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_e2e.py -m gpu -q > gpurun_out/pytest_e2e.txt 2>&1
tail -15 gpurun_out/pytest_e2e.txt
