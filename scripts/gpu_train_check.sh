#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests_all.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/gpu_tests_all.log
tail -5 gpurun_out/gpu_tests_all.log
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_b128.json 2> gpurun_out/bench_train_b128.err
cat gpurun_out/bench_train_b128.json; tail -3 gpurun_out/bench_train_b128.err
timeout 900 python bench.py --mode train --math bf16x3 --batch 64 --steps 5 --warmup 3 > gpurun_out/bench_train_b64_x3.json 2> gpurun_out/bench_train_b64_x3.err
cat gpurun_out/bench_train_b64_x3.json; tail -3 gpurun_out/bench_train_b64_x3.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
