#!/bin/bash
# round-2 evidence run: driver-style test suite + smoke, default bench line, ncu launch list + --set full captures
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r02}
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name: $*" | tee gpurun_out/$name.log
  timeout "$to" "$@" >> gpurun_out/$name.log 2>&1
  echo "=== $name exit $?" | tee -a gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run b_tests 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
run b_smoke 600 python __graft_entry__.py smoke
run b_bench 1200 python bench.py --steps 10 --warmup 3
cp gpurun_out/b_bench.log gpurun_out/${TAG}_bench_default.log
run b_memcheck 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_forward.py
bash scripts/gpu_prof.sh ${TAG} 256
for f in b_tests b_smoke b_bench b_memcheck; do echo "----- $f"; tail -n ${TAILN:-12} gpurun_out/$f.log; done
