#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged or standalone or small_strict or brush or depth_stencil or multi_model" > gpurun_out/san_race.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/san_race.txt | tail -3
grep -E "Warning: Race" gpurun_out/san_race.txt | sed 's/0x[0-9a-f]*//g' | cut -c1-200 | sort | uniq -c | sort -rn | head -5
python scripts/strips_8k.py > gpurun_out/strips_n1.txt 2>&1; tail -1 gpurun_out/strips_n1.txt | cut -c1-250
python scripts/stage_times.py --n 6000000 --cams outside 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cam'], round(d['frame_ms'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})"
