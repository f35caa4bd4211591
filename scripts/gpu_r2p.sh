#!/bin/bash
# round-2 profiling session: re-run the two fixed test files, test durations, then ncu (launch list + --set full)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name: $*" | tee gpurun_out/$name.log
  timeout "$to" "$@" >> gpurun_out/$name.log 2>&1
  echo "=== $name exit $?" | tee -a gpurun_out/$name.log; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -m gpu"
run p_fix   900 $PT tests/test_optim.py tests/test_gpu_backward.py -k "adamw or fixtures or batch8" --durations=8
run p_dur   900 $PT tests/test_gpu_kernels.py -k "f16c and not attention" --durations=12
bash scripts/gpu_prof.sh r02a 256
for f in gemm attn fuse; do ncu -i gpurun_out/r02a_$f.ncu-rep --page raw --csv > gpurun_out/r02a_${f}_raw.csv 2>/dev/null; done
for f in p_fix p_dur; do echo "----- $f"; tail -n 25 gpurun_out/$f.log; done
