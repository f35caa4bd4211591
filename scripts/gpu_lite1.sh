#!/bin/bash
# Lite (config 5) class breakdown, fused vs split MLP
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for T in 27 243; do
 for fl in 0 0x100; do
  timeout 300 python bench.py --model lite --batch 512 --frames $T --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-flags $fl 2>gpurun_out/lite1_${T}_${fl}.err > gpurun_out/lite1_${T}_${fl}.log
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/lite1_${T}_${fl}.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("T=$T flags $fl:", round(j["value"], 1), "seq/s", round(j["ms_per_step"], 2), "ms", j["clocks"], 
          {k: round(v, 2) for k, v in r.get("class_ms_per_step", {}).items()}, (j.get("parity") or {}).get("ok"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/lite1_${T}_${fl}.err").read()[-1500:])
PY
 done
done
