#!/bin/bash
# round-2 two-GPU evidence (run with `gpurun --gpus 2`): the multi-GPU tests a 1-GPU box skips, and the N=2 bench line
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name: $*" | tee gpurun_out/$name.log
  timeout "$to" "$@" >> gpurun_out/$name.log 2>&1
  echo "=== $name exit $?" | tee -a gpurun_out/$name.log; }
run c_tests2 1200 python -m pytest tests/test_gpu_ddp.py tests/test_gpu_forward.py -q -m gpu -k "ddp or two_gpu or data_parallel" -p no:cacheprovider -rs
run c_bench2 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3
for f in c_tests2 c_bench2; do echo "----- $f"; tail -n 12 gpurun_out/$f.log; done
