#!/bin/bash
# round-2 profiling session: ncu launch list + --set full captures; only CSV exports come back (reports exceed the 64 MiB cap)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r02a}; B=${2:-256}
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv \
   python scripts/prof_forward.py --batch $B > gpurun_out/${TAG}_launches.log 2>&1
cap() {  # name, kernel regex, count
  timeout 900 $NCU --set full --import-source on -k regex:$2 -c $3 -o /tmp/${TAG}_$1 -f \
     python scripts/prof_forward.py --batch $B > gpurun_out/${TAG}_$1.log 2>&1
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_$1.ncu-rep --page source --csv > gpurun_out/${TAG}_$1_source.csv 2>/dev/null
  ls -la /tmp/${TAG}_$1.ncu-rep gpurun_out/${TAG}_$1_raw.csv gpurun_out/${TAG}_$1_source.csv
}
cap gemm "gemm2_kernel|mlp_fused" 4
cap attn attn_ 2
cap fuse fuse_kernel 1
du -sh gpurun_out; tail -n 3 gpurun_out/${TAG}_gemm.log gpurun_out/${TAG}_attn.log
