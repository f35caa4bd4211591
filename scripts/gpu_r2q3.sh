#!/bin/bash
# A/B on ONE box: 16-warp F16C epilogues (default) vs 8-warp (flag 0x80); quick correctness pass first
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name: $*" | tee gpurun_out/$name.log
  timeout "$to" "$@" >> gpurun_out/$name.log 2>&1
  echo "=== $name exit $?" | tee -a gpurun_out/$name.log; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu -x"
run q_kern  900 $PT tests/test_gpu_kernels.py -k "f16c"
run q_fwd   900 $PT tests/test_gpu_forward.py
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras"
run q_ab_a1 600 $B
run q_ab_b1 600 $B --kernel-flags 0x80
run q_ab_a2 600 $B
run q_ab_l1 600 $B --model lite --batch 512
run q_ab_l2 600 $B --model lite --batch 512 --kernel-flags 0x80
for f in q_kern q_fwd; do echo "----- $f"; tail -n 4 gpurun_out/$f.log; done
for f in q_ab_a1 q_ab_b1 q_ab_a2 q_ab_l1 q_ab_l2; do echo "----- $f"; tail -n 2 gpurun_out/$f.log | cut -c1-250; done
