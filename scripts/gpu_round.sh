#!/bin/bash
# new tests + config-3 bench + per-kernel time list of one training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_losses_tta.py tests/test_gpu_backward.py -q -m gpu > gpurun_out/new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/new_tests.log; tail -12 gpurun_out/new_tests.log
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_b128.json 2> gpurun_out/bench_train_b128.err
cat gpurun_out/bench_train_b128.json; tail -3 gpurun_out/bench_train_b128.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
    python scripts/bench_train.py --batch 32 --frames 243 --steps 1 --warmup 1 --skip-torch > gpurun_out/train_ncu.log 2>&1
python scripts/ncu_train_summary.py gpurun_out/train_launches.csv gpurun_out/train_step_b32_kernels.json
