#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6
for rep in 1 2; do
for lib in default ab_libs/old.so; do
  if [ $lib = default ]; then unset MB_LIB_OVERRIDE; else export MB_LIB_OVERRIDE=$PWD/$lib; fi
  timeout 600 python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('train $lib:', round(j['value'],1), j['unit'], round(j['ms_per_step'],1),'ms', j['clocks']['sm_mhz'])"
done; done
