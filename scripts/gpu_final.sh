#!/bin/bash
# End-of-round validation exactly as the driver runs it: the whole -m gpu suite in one process, smoke(), both bench arms.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/final_gpu_tests.log; tail -4 gpurun_out/final_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/final_smoke.log; tail -5 gpurun_out/final_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
cut -c1-500 gpurun_out/final_bench_ref.json; tail -2 gpurun_out/final_bench_ref.err
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/final_bench.json"))
r = j["roofline"]
print("bench:", round(j["value"], 1), j["unit"], round(j["ms_per_step"], 1), "ms/step; e2e", round(j["e2e"]["value"], 1),
      "; gemm frac", round(r["frac"], 3), "executed", round(r["tensor_pipe_tflops_executed"], 1), "TF/s; launches", j["gpu_launches"],
      "; clocks", j["clocks"], "; cpu", j.get("cpu_baseline", {}).get("value"))
print({k: round(v, 1) for k, v in r["class_ms_per_step"].items()})
PY
tail -2 gpurun_out/final_bench.err
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/final_bench_train.json 2> gpurun_out/final_bench_train.err
cut -c1-330 gpurun_out/final_bench_train.json; tail -2 gpurun_out/final_bench_train.err
