#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear" -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
bash scripts/gpu_ab_lib.sh ${1:-r02j_ab}
