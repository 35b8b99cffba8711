#!/bin/bash
# strip-cull tests + the bench (incl. the 8K strip frame) at 1 and 2 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -k "strip" -x -q > gpurun_out/pytest_strips.txt 2>&1
tail -5 gpurun_out/pytest_strips.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || tail -5 gpurun_out/bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("N=1", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 3) for k, v in d["stages"]["ms"].items()})
        print("strips", json.dumps(d.get("strips"))[:1500])
    except Exception as e:
        print("bench parse failed", e)
PY
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -20 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
    print("N=" + n, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
    print("strips", json.dumps(d.get("strips"))[:2500])
except Exception as e:
    print("bench parse failed", e)
PY
