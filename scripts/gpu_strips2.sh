#!/bin/bash
# 2-GPU strip test + strips timing only (scripts/strips_8k.py), both preprocessors
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_parity.py -k "strips_over_two" -x -q 2>&1 | tail -5
for pc in 0 1; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/strips_8k.py --check --partition-cull $pc 2>&1 | grep '^{' | cut -c1-600
done
