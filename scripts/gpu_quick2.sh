#!/bin/bash
# quick loop for GEMM changes: 2-CTA linear kernel tests + forward tests + base bench + Lite bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout 300 -m gpu"
timeout 600 $PT tests/test_gpu_kernels.py -k "tc2cta or single_pass" > gpurun_out/q_gemm.log 2>&1; echo "gemm exit $?"; tail -2 gpurun_out/q_gemm.log
timeout 900 $PT tests/test_gpu_forward.py -k "not simt" > gpurun_out/q_fwd.log 2>&1; echo "fwd exit $?"; tail -2 gpurun_out/q_fwd.log
for cfg in "--model base" "--model lite --batch 512" "--model lite --batch 512 --frames 27 --steps 20"; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $cfg > gpurun_out/bench_q.log 2>&1; echo "bench [$cfg] exit $?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_q.log') if x.startswith('{')]
if not l: print(open('gpurun_out/bench_q.log').read()[-1500:])
else:
    d=json.loads(l[-1]); r=d['roofline']
    print(' ', round(d['value'],1), 'seq/s', round(d['ms_per_step'],2), 'ms  clk', d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1), ' gemm frac', round(r['frac'],3))
    print(' ', {k: round(v,2) for k,v in r['class_ms_per_step'].items()})
PY
done
