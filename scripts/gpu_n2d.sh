#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_ddp.py tests/test_gpu_forward.py -q -m gpu -k "ddp or data_parallel or gradient_exchange or two_gpu" --tb=short 2>&1 | grep -v CUDAGuardImpl | tail -60 | tee gpurun_out/n2d_tests.log
