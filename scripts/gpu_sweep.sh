#!/bin/bash
# Extra bench points (BASELINE configs 2/5 variants): results in gpurun_out/sweep.jsonl
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
run() { echo "## $*" >> gpurun_out/sweep.jsonl; timeout 600 python bench.py --no-cpu-baseline "$@" 2>>gpurun_out/sweep.err | tail -1 >> gpurun_out/sweep.jsonl; }
run --model base --batch 256 --frames 243 --math bf16 --steps 5
run --model lite --batch 512 --frames 243 --steps 5
run --model lite --batch 512 --frames 81 --steps 10
run --model lite --batch 512 --frames 27 --steps 20
run --model lite --batch 512 --frames 243 --math bf16 --steps 5
run --model lite --batch 512 --frames 27 --math bf16 --steps 20
run --model base --batch 1 --frames 243 --steps 20
run --model lite --batch 1 --frames 27 --steps 50
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    if l.startswith('##'): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception as e: print('ERR', l[:200]); continue
    r=d['roofline']
    print(f"  value {d['value']:.1f} seq/s  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.1f}  gemm frac {r['frac']:.3f} ({r['achieved']:.0f} TF/s alg, passes {r['mma_passes']})  whole-step frac {r['whole_step_frac']:.3f}  clk {d['clocks']['sm_mhz']}")
    print("   ", {k: round(v,2) for k,v in r['class_ms_per_step'].items()})
PY
