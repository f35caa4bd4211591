#!/bin/bash
# cycle-level timing experiments: ncu (clocks not locked) counts SM cycles / tensor-pipe activity of the first 8 GEMM launches of
# one forward for the default library and for every experiment build in ab_libs/ (power-cap clock changes cancel out in cycles)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-exp}
M=sm__cycles_elapsed.max,gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum
for lib in default $(ls ab_libs/*.so 2>/dev/null); do
  name=$(basename $lib .so)
  if [ $lib = default ]; then unset MB_LIB_OVERRIDE; else export MB_LIB_OVERRIDE=$PWD/$lib; fi
  timeout 600 ncu --clock-control none --profile-from-start off -k regex:gemm2_kernel -c 8 --metrics $M --csv \
     --log-file gpurun_out/${TAG}_${name}.csv python scripts/prof_forward.py --batch 256 > gpurun_out/${TAG}_${name}.log 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_${name}.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
out = {}
for r in rows[1:]:
    out.setdefault((r[ii], r[ki].split("(")[0][-28:]), {})[r[mi].split(".")[0][-28:]] = r[vi]
print("== $name")
for k, v in list(out.items())[:4]:
    print("  ", k[1], v)
PY
done
