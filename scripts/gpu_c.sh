#!/bin/bash
for A in 0 1 2 3; do echo "ablate=$A"; SB_SORT_ABLATE=$A python scripts/sort_bench.py --n 64000000 --iters 5 2>&1 | grep u32; done
