#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention_core_backward" > gpurun_out/abw_tests.log 2>&1
echo "abw tests exit $?" >> gpurun_out/abw_tests.log; tail -6 gpurun_out/abw_tests.log
timeout 900 python -m pytest tests/test_gpu_backward.py -q -m gpu > gpurun_out/bwd_tests.log 2>&1
echo "backward tests exit $?" >> gpurun_out/bwd_tests.log; tail -4 gpurun_out/bwd_tests.log
timeout 600 python scripts/bench_train.py --batch 32 --frames 243 --skip-torch | tee gpurun_out/train_base.json
timeout 600 python scripts/bench_train.py --batch 64 --frames 243 --lite --skip-torch | tee gpurun_out/train_lite.json
timeout 600 python scripts/bench_train.py --batch 256 --frames 27 --lite --skip-torch | tee gpurun_out/train_lite_t27.json
