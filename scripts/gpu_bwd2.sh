#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_losses_tta.py tests/test_gpu_forward.py -q -s -m gpu -k "backward or loss or tta or drop_path or autograd" > gpurun_out/bwd_tests.log 2>&1
echo "backward tests exit $?" >> gpurun_out/bwd_tests.log; grep -v "per-class" gpurun_out/bwd_tests.log | grep -v "^    " | tail -15
timeout 600 python scripts/bench_train.py --batch 32 --frames 243 --skip-torch | tee gpurun_out/train_base.json
