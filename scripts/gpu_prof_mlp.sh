#!/bin/bash
# ncu --set full (+source) of the fused MLP kernel: gpu_prof_mlp.sh <tag> [batch] [model]
cd "$(dirname "$0")/.." || exit 1
TAG=${1:-r02h}; B=${2:-32}; MODEL=${3:-base}
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --set full --import-source on -k regex:mlp_fused -c 1 -o gpurun_out/${TAG}_mlp_${MODEL} -f \
   python scripts/prof_forward.py --batch $B --model $MODEL > gpurun_out/${TAG}_mlp_${MODEL}.log 2>&1
tail -3 gpurun_out/${TAG}_mlp_${MODEL}.log; ls -la gpurun_out | tail -5
