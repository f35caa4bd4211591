#!/bin/bash
# same-box A/B of kernel flag sets: gpu_ab_flags.sh <tag> <flags1> <flags2> ...   (bench forward only, two rounds)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=$1; shift
for rep in 1 2; do
  for fl in "$@"; do
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --kernel-flags $fl 2>gpurun_out/${TAG}_${fl}_$rep.err > gpurun_out/${TAG}_${fl}_$rep.log
    python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_${fl}_$rep.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("flags $fl $rep:", round(j["value"], 1), "seq/s", round(j["ms_per_step"], 1), "ms", j["clocks"]["sm_mhz"], "MHz",
          {k: round(v, 1) for k, v in r.get("class_ms_per_step", {}).items()}, (j.get("parity") or {}).get("ok"))
except Exception as e:
    print("flags $fl $rep: FAILED", e); print(open("gpurun_out/${TAG}_${fl}_$rep.err").read()[-1500:])
PY
  done
done
