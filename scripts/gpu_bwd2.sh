#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -s -m gpu > gpurun_out/bwd_tests.log 2>&1
echo "backward tests exit $?" >> gpurun_out/bwd_tests.log; grep -v "per-class" gpurun_out/bwd_tests.log | grep -v "^    " | tail -12
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_b128.json 2> gpurun_out/bench_train_b128.err
cut -c1-330 gpurun_out/bench_train_b128.json; tail -2 gpurun_out/bench_train_b128.err
nvidia-smi --query-gpu=memory.used --format=csv
