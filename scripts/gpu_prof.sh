#!/bin/bash
# ncu evidence: launch list of one forward + full captures of the hot kernels.  Usage: gpu_prof.sh <tag> [batch]
cd "$(dirname "$0")/.." || exit 1
TAG=${1:-r01}; B=${2:-32}
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv \
   python scripts/prof_forward.py --batch $B > gpurun_out/${TAG}_launches.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:gemm2_kernel -c 4 -o gpurun_out/${TAG}_gemm -f \
   python scripts/prof_forward.py --batch $B > gpurun_out/${TAG}_gemm.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:attn_ -c 2 -o gpurun_out/${TAG}_attn -f \
   python scripts/prof_forward.py --batch $B > gpurun_out/${TAG}_attn.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:fuse_kernel -c 1 -o gpurun_out/${TAG}_fuse -f \
   python scripts/prof_forward.py --batch $B > gpurun_out/${TAG}_fuse.log 2>&1
ls -la gpurun_out/ | tail -20
tail -3 gpurun_out/${TAG}_gemm.log gpurun_out/${TAG}_attn.log
