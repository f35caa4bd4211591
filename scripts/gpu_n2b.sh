#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -m gpu > gpurun_out/ddp_test.log 2>&1; tail -5 gpurun_out/ddp_test.log
