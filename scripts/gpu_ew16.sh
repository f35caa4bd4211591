#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -x -m gpu -k "backward or wgrad or bf16 or linear or golden" > gpurun_out/ew16_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/ew16_tests.log; tail -6 gpurun_out/ew16_tests.log
timeout 600 python bench.py --math bf16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fwd_bf16.json 2> gpurun_out/bench_fwd_bf16.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_fwd_bf16.json"))
print("fwd bf16:", round(j["value"], 1), "seq/s", round(j["ms_per_step"], 1), "ms", {k: round(v, 1) for k, v in j["roofline"]["class_ms_per_step"].items()})
PY
timeout 600 python scripts/bench_train.py --batch 32 --frames 243 --skip-torch | tee gpurun_out/train_base.json
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_b128.json 2> gpurun_out/bench_train_b128.err
cut -c1-330 gpurun_out/bench_train_b128.json; tail -2 gpurun_out/bench_train_b128.err
