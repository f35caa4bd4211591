#!/bin/bash
# GELU (one MUFU) + packed LN statistics: tests touching the changed epilogues, then a same-box A/B against ab_libs/prev.so;
# Lite T=27 A/B as well
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_mlp_fused.py tests/test_gpu_forward.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2j_tests.log
if grep -q "failed\|error" gpurun_out/r2j_tests.log; then echo "TESTS FAILED"; fi
bash scripts/gpu_ab_lib.sh r2j
for lib in default ab_libs/prev.so; do
  if [ $lib = default ]; then unset MB_LIB_OVERRIDE; else export MB_LIB_OVERRIDE=$PWD/$lib; fi
  timeout 300 python bench.py --model lite --batch 512 --frames 27 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lite T=27 $lib', round(j['value'],1), 'seq/s', j['clocks']['sm_mhz'], {k: round(v,2) for k,v in j['roofline']['class_ms_per_step'].items()})"
done
