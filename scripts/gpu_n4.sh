#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err
cat gpurun_out/bench_train_n$N.json | cut -c1-400; tail -2 gpurun_out/bench_train_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_fwd_n$N.json 2> gpurun_out/bench_fwd_n$N.err
cat gpurun_out/bench_fwd_n$N.json | cut -c1-400; tail -2 gpurun_out/bench_fwd_n$N.err
