#!/bin/bash
# two-GPU evidence of the current build: DataParallel + NCCL gradient-exchange tests, config-4 style step, forward at N=2
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/n2c_gpus.txt
timeout 600 python -m pytest tests/test_gpu_ddp.py tests/test_gpu_forward.py -q -m gpu -k "ddp or data_parallel or gradient_exchange or two_gpu" --tb=short 2>&1 | grep -v CUDAGuardImpl | tail -40 | tee gpurun_out/n2c_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/n2c_bench_train.json 2> gpurun_out/n2c_bench_train.err
cut -c1-700 gpurun_out/n2c_bench_train.json; tail -2 gpurun_out/n2c_bench_train.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/n2c_bench_fwd.json 2> gpurun_out/n2c_bench_fwd.err
cut -c1-400 gpurun_out/n2c_bench_fwd.json; tail -2 gpurun_out/n2c_bench_fwd.err
