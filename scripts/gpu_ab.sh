#!/bin/bash
# A/B on ONE box: forward bench of the current tree vs the forward-only tree in .ab_old (same GPU, alternating)
mkdir -p gpurun_out
for rep in 1 2; do
  for which in new old; do
    if [ $which = new ]; then dir=.; else dir=.ab_old; fi
    (cd $dir && timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null) > gpurun_out/ab_${which}_$rep.json
    python - <<PY
import json
j = json.load(open("gpurun_out/ab_${which}_$rep.json"))
r = j["roofline"]
print("$which $rep:", round(j["value"], 1), "seq/s", round(j["ms_per_step"], 1), "ms", j["clocks"]["sm_mhz"], "MHz", {k: round(v, 1) for k, v in r["class_ms_per_step"].items()})
PY
  done
done
nvidia-smi --query-gpu=name,power.limit,temperature.gpu,clocks.sm --format=csv
