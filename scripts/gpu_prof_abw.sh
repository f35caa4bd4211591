#!/bin/bash
# ncu of the attention backward kernels (temporal q / kv), with source-level stall sampling
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --set full --import-source on -k regex:attn_bwd -c 2 -o gpurun_out/abw -f \
   python scripts/prof_train.py --batch 32 > gpurun_out/abw.log 2>&1
ncu -i gpurun_out/abw.ncu-rep --page raw --csv > gpurun_out/abw_raw.csv 2>/dev/null
python scripts/ncu_compact.py gpurun_out/abw_raw.csv
echo "---- q kernel top stalls"; python scripts/ncu_top_stalls.py gpurun_out/abw.ncu-rep attn_bwd_q 0 28
echo "---- kv kernel top stalls"; python scripts/ncu_top_stalls.py gpurun_out/abw.ncu-rep attn_bwd_kv 0 22
rm -f gpurun_out/abw.ncu-rep
