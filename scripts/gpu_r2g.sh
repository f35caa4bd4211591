#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear or gemm" -p no:cacheprovider 2>&1 | tail -5
bash scripts/gpu_exp_cycles.sh r02g
bash scripts/gpu_ab_lib.sh r02g_ab
