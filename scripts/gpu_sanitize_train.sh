#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_train.py > gpurun_out/sanitize_train.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitize_train.log; tail -12 gpurun_out/sanitize_train.log
timeout 900 python -m pytest tests/test_gpu_backward.py -q -s -m gpu > gpurun_out/bwd_tests.log 2>&1
echo "backward tests exit $?" >> gpurun_out/bwd_tests.log; grep -v "per-class" gpurun_out/bwd_tests.log | tail -30
