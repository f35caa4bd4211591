#!/bin/bash
# native backward: gradient parity + kernel-level backward hooks, output kept in gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -s -m gpu > gpurun_out/bwd_tests.log 2>&1
echo "backward tests exit $?" >> gpurun_out/bwd_tests.log
tail -80 gpurun_out/bwd_tests.log
