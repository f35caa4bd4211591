#!/bin/bash
# Run on the GPU box (via gpurun): isolated pytest processes per kernel family so that a trap in one
# family does not poison the CUDA context of the others; logs land in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, cmd...
  local name=$1; shift; local to=$1; shift
  echo "=== $name: $*" | tee gpurun_out/$name.log
  timeout "$to" "$@" >> gpurun_out/$name.log 2>&1
  echo "=== $name exit $?" | tee -a gpurun_out/$name.log
}
PT="python -m pytest -q -p no:cacheprovider --timeout 300 -m gpu"
run k_ref      600 $PT tests/test_gpu_kernels.py -k "simt_ref" -x
run k_gemm     600 $PT tests/test_gpu_kernels.py -k "tc2cta and linear_bf16x3"
run k_gemm_old 600 $PT tests/test_gpu_kernels.py -k "tc1cta and linear_bf16x3"
run k_gemm1    300 $PT tests/test_gpu_kernels.py -k "single_pass and linear"
run k_attn     600 $PT tests/test_gpu_kernels.py -k "tcgen05 and temporal_attention"
run k_attns    600 $PT tests/test_gpu_kernels.py -k "tcgen05 and spatial_attention"
run k_attn1    300 $PT tests/test_gpu_kernels.py -k "temporal_attention_bf16"
run f_simt     900 $PT tests/test_gpu_forward.py -k "simt" -s
run f_main     900 $PT tests/test_gpu_forward.py -k "not simt" -s
run smoke      600 python __graft_entry__.py smoke
[ -n "$SKIP_BENCH_S" ] || run bench_s    600 python bench.py --steps 3 --warmup 3 --batch 32 --no-cpu-baseline
run bench      900 python bench.py --steps 5 --warmup 3
for f in k_ref k_gemm k_gemm_old k_gemm1 k_attn k_attns k_attn1 f_simt f_main smoke bench_s bench; do
  echo "----- $f"; tail -n ${TAILN:-12} gpurun_out/$f.log
done
