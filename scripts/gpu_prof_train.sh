#!/bin/bash
# ncu --set full of the training step's tensor-core kernels (bf16 mode).  Usage: gpu_prof_train.sh <tag> [batch]
cd "$(dirname "$0")/.." || exit 1
TAG=${1:-r01t}; B=${2:-64}
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
# depth-0 of the forward (first 12 gemm launches) ...
timeout 900 $NCU --set full --import-source on -k regex:gemm2_kernel -c 10 -o gpurun_out/${TAG}_gemm_fwd -f \
   python scripts/prof_train.py --batch $B > gpurun_out/${TAG}_gemm_fwd.log 2>&1
# ... and the first sublayers of the backward (launch 81.. = after the 81 forward GEMMs)
timeout 900 $NCU --set full --import-source on -k regex:gemm2_kernel -s 82 -c 14 -o gpurun_out/${TAG}_gemm_bwd -f \
   python scripts/prof_train.py --batch $B > gpurun_out/${TAG}_gemm_bwd.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:wgrad_kernel -s 1 -c 6 -o gpurun_out/${TAG}_wgrad -f \
   python scripts/prof_train.py --batch $B > gpurun_out/${TAG}_wgrad.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:attn_bwd -c 4 -o gpurun_out/${TAG}_attnbwd -f \
   python scripts/prof_train.py --batch $B > gpurun_out/${TAG}_attnbwd.log 2>&1
for part in gemm_fwd gemm_bwd wgrad attnbwd; do
  ncu -i gpurun_out/${TAG}_${part}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${part}_raw.csv 2>/dev/null
done
python scripts/ncu_compact.py gpurun_out/${TAG}_gemm_fwd_raw.csv gpurun_out/${TAG}_gemm_bwd_raw.csv gpurun_out/${TAG}_wgrad_raw.csv gpurun_out/${TAG}_attnbwd_raw.csv --json gpurun_out/${TAG}_train_ncu_summary.json 2>&1 | tail -60
ls -la gpurun_out | tail -12
