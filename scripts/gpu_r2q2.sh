#!/bin/bash
# A/B on ONE box: (a) default build, (b) MMA order variant, (c) polling-wait build; plus the quick correctness pass
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
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DMB_BUILD -DMB_WAIT_HINT_NS=0u \
   -o /tmp/libmb_poll.so motionbert_b200/csrc/mb_api.cu > gpurun_out/q_pollbuild.log 2>&1
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras"
run q_ab_a1 600 $B
run q_ab_b1 600 $B --kernel-flags 0x80
MB_LIB_OVERRIDE=/tmp/libmb_poll.so run q_ab_c1 600 $B
run q_ab_a2 600 $B
run q_ab_b2 600 $B --kernel-flags 0x80
for f in q_kern q_fwd q_bwd; do echo "----- $f"; tail -n 4 gpurun_out/$f.log; done
for f in q_ab_a1 q_ab_b1 q_ab_c1 q_ab_a2 q_ab_b2; do echo "----- $f"; tail -n 2 gpurun_out/$f.log | cut -c1-250; done
