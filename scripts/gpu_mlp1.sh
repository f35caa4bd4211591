#!/bin/bash
# fused MLP bring-up: bit-exactness tests, then a same-box A/B (fused ring / split / fused no-ring) of the forward bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_mlp_fused.py -x -q 2>&1 | tail -25 | tee gpurun_out/mlp1_tests.log
if grep -q "failed\|error\|Error" gpurun_out/mlp1_tests.log; then echo "TESTS FAILED - skipping bench"; exit 1; fi
bash scripts/gpu_ab_flags.sh ${1:-mlp2} 0 0x800 0x400 0x100
