#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_losses_tta.py -q -s -m gpu > gpurun_out/bwd_tests.log 2>&1
echo "backward tests exit $?" >> gpurun_out/bwd_tests.log; grep -v "per-class" gpurun_out/bwd_tests.log | tail -8
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 600 python scripts/bench_train.py --batch 32 --frames 243 --skip-torch | tee gpurun_out/train_base.json
timeout 600 python scripts/bench_train.py --batch 64 --frames 243 --lite --skip-torch | tee gpurun_out/train_lite.json
