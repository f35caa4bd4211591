#!/bin/bash
# round-2 final evidence of the current build: the driver's sequence (whole -m gpu suite, smoke, both bench arms), memcheck
# of the forward, then the ncu launch list + --set full captures at the bench configuration
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r02k}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt
bash scripts/gpu_final.sh 2>&1 | tee gpurun_out/${TAG}_final.log
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_forward.py > gpurun_out/${TAG}_memcheck.log 2>&1
echo "=== memcheck exit $?" >> gpurun_out/${TAG}_memcheck.log; tail -4 gpurun_out/${TAG}_memcheck.log
bash scripts/gpu_r2p.sh ${TAG} 256 2>&1 | tail -12
