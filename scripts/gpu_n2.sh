#!/bin/bash
# two-GPU runs: config 4 style training step (gradient all-reduce over NCCL) and the headline forward
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err
cat gpurun_out/bench_train_n2.json; tail -3 gpurun_out/bench_train_n2.err
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu -k "data_parallel" > gpurun_out/dp_test.log 2>&1; tail -3 gpurun_out/dp_test.log
timeout 900 python -m pytest tests/test_gpu_ddp.py -q -m gpu > gpurun_out/ddp_test.log 2>&1; tail -5 gpurun_out/ddp_test.log
