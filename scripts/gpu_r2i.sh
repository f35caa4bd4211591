#!/bin/bash
# round-2 check of the final fused-MLP kernel + auto-graph: tests, latency, A/B
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_mlp_fused.py tests/test_gpu_forward.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2i_tests.log
timeout 300 python scripts/latency_graph.py 2>&1 | tail -8 | tee gpurun_out/r2i_latency.log
