#!/bin/bash
# round-2 evidence run: driver-style test suite + smoke, default bench line, reference arm, memcheck, then the ncu session
# (launch list + --set full captures, CSV exports only: the reports themselves exceed gpurun's 64 MiB copy-back cap)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r02}
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name: $*" | tee gpurun_out/$name.log
  timeout "$to" "$@" >> gpurun_out/$name.log 2>&1
  echo "=== $name exit $?" | tee -a gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
run ${TAG}_tests 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
run ${TAG}_smoke 600 python __graft_entry__.py smoke
run ${TAG}_bench_default 1200 python bench.py --steps 10 --warmup 3
run ${TAG}_bench_reference_arm 900 python bench.py --impl reference --steps 3 --warmup 1
run ${TAG}_memcheck 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_forward.py
bash scripts/gpu_r2p.sh ${TAG} 256
for f in tests smoke bench_default bench_reference_arm memcheck; do echo "----- $f"; tail -n ${TAILN:-6} gpurun_out/${TAG}_$f.log | cut -c1-1500; done
