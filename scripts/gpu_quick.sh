#!/bin/bash
# quick loop: temporal/spatial attention kernel tests + forward tests + bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 -m gpu"
timeout 600 $PT tests/test_gpu_kernels.py -k "attention" > gpurun_out/q_attn.log 2>&1; echo "attn exit $?"; tail -3 gpurun_out/q_attn.log
timeout 900 $PT tests/test_gpu_forward.py -k "not simt" > gpurun_out/q_fwd.log 2>&1; echo "fwd exit $?"; tail -3 gpurun_out/q_fwd.log
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; echo "bench exit $?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')]
if not l: print(open('gpurun_out/bench.log').read()[-1500:])
else:
    d=json.loads(l[-1]); r=d['roofline']
    print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], 'e2e', d['e2e']['value'])
    print({k: round(v,2) for k,v in r['class_ms_per_step'].items()})
    print('gemm frac', r['frac'], 'executed TF/s', r['tensor_pipe_tflops_executed'])
PY
