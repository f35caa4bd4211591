#!/bin/bash
# same-box A/B of alternative builds of the product library: ab_libs/<name>.so are loaded through MB_LIB_OVERRIDE
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-ab}
for rep in 1 2; do
  for lib in default $(ls ab_libs/*.so 2>/dev/null); do
    name=$(basename $lib .so)
    if [ $lib = default ]; then unset MB_LIB_OVERRIDE; else export MB_LIB_OVERRIDE=$PWD/$lib; fi
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/${TAG}_${name}_$rep.err > gpurun_out/${TAG}_${name}_$rep.log
    python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_${name}_$rep.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("$name $rep:", round(j["value"], 1), "seq/s", round(j["ms_per_step"], 1), "ms", j["clocks"]["sm_mhz"], "MHz",
          {k: round(v, 1) for k, v in r.get("class_ms_per_step", {}).items()}, (j.get("parity") or {}).get("ok"))
except Exception as e:
    print("$name $rep: FAILED", e); print(open("gpurun_out/${TAG}_${name}_$rep.err").read()[-1500:])
PY
  done
done
