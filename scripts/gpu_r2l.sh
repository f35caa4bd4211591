#!/bin/bash
# L2 policy of the fused MLP kernel: bit-exactness tests, DRAM bytes per launch (ncu metrics) with / without the hints, A/B bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_mlp_fused.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2l_tests.log
for fl in 0 0x800; do
  timeout 600 ncu --clock-control none --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_op_write_hit_rate.pct,lts__t_sector_op_read_hit_rate.pct \
     -k regex:mlp_fused -c 2 --csv --log-file gpurun_out/r2l_dram_$fl.csv python scripts/prof_forward.py --batch 256 --kernel-flags $fl > gpurun_out/r2l_dram_$fl.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2l_dram_$fl.csv")) if len(r)>10]
h=rows[0]; n=h.index("Metric Name"); v=h.index("Metric Value"); u=h.index("Metric Unit")
print("flags $fl:", [(r[n].split("__")[-1][:28], r[v], r[u]) for r in rows[1:]])
PY
done
bash scripts/gpu_ab_flags.sh r2l 0 0x800
