#!/bin/bash
# usage: gpu_mg.sh <N>
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['single_frame'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 scripts/strips_8k.py --check > gpurun_out/strips_n$N.txt 2>&1
tail -2 gpurun_out/strips_n$N.txt | cut -c1-300
