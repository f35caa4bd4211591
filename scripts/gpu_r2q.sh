#!/bin/bash
# quick validation + A/B after an optimisation step: kernel tests, golden forward, backward, short bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name: $*" | tee gpurun_out/$name.log
  timeout "$to" "$@" >> gpurun_out/$name.log 2>&1
  echo "=== $name exit $?" | tee -a gpurun_out/$name.log; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu -x"
run q_kern  900 $PT tests/test_gpu_kernels.py
run q_fwd   900 $PT tests/test_gpu_forward.py
run q_bwd   900 $PT tests/test_gpu_backward.py
run q_bench 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras
for f in q_kern q_fwd q_bwd; do echo "----- $f"; tail -n 6 gpurun_out/$f.log; done
tail -n 3 gpurun_out/q_bench.log | cut -c1-1800
