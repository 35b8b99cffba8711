#!/bin/bash
# A/B of runtime switches: per-stage times of the 6 M bench scene (outside + inside camera) under each environment setting.
# Usage: gpu_ab.sh [--tests] [--size W H] "ENV1=a ENV2=b" "" ...      ("" = defaults)
mkdir -p gpurun_out
TESTS=0; SIZE="1920 1080"; MODE=0
while [[ "$1" == --* ]]; do
  case "$1" in
    --tests) TESTS=1; shift;;
    --size) SIZE="$2 $3"; shift 3;;
    --mode) MODE="$2"; shift 2;;
  esac
done
for E in "$@"; do
  echo "== [$E] size $SIZE mode $MODE"
  env $E python scripts/stage_times.py --n 6000000 --cams outside inside --size $SIZE --mode $MODE 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cam'], 'frame', round(d['frame_ms'],3), 'D', d['duplicates'], {k:round(v,4) for k,v in d['stages_ms'].items()})"
done
if [ $TESTS = 1 ]; then timeout 1400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
