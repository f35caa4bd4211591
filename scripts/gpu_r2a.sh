#!/bin/bash
# round-2 first check: F16C GEMM unit tests, golden forward in both modes, A/B bench f16c vs bf16x3, backward fixtures
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name: $*" | tee gpurun_out/$name.log
  timeout "$to" "$@" >> gpurun_out/$name.log 2>&1
  echo "=== $name exit $?" | tee -a gpurun_out/$name.log; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu"
run a_f16c   600 $PT tests/test_gpu_kernels.py -k "f16c and not attention" -s
run a_attn16 600 $PT tests/test_gpu_kernels.py -k "attention_f16c" -s
run a_gold   900 $PT tests/test_gpu_forward.py -k "golden and not simt" -s
run a_bench_f16c 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run a_bench_x3   600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --math bf16x3
run a_fwd    900 $PT tests/test_gpu_forward.py -k "not golden" -s
run a_misc   900 $PT tests/test_action.py tests/test_optim.py tests/test_augment.py tests/test_gpu_losses_tta.py -s
run a_bwd    1500 $PT tests/test_gpu_backward.py -s
for f in a_f16c a_attn16 a_gold a_bench_f16c a_bench_x3 a_fwd a_misc a_bwd; do echo "----- $f"; tail -n ${TAILN:-15} gpurun_out/$f.log; done
