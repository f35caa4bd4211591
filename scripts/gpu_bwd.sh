#!/bin/bash
# native backward: gradient parity + training-step timing + per-kernel time list, output kept in gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -s -m gpu > gpurun_out/bwd_tests.log 2>&1
echo "backward tests exit $?" >> gpurun_out/bwd_tests.log
grep -v "per-class median" gpurun_out/bwd_tests.log | tail -60
timeout 600 python scripts/bench_train.py --batch 32 --frames 243 > gpurun_out/train_base.json 2> gpurun_out/train_base.err; cat gpurun_out/train_base.json; tail -3 gpurun_out/train_base.err
timeout 600 python scripts/bench_train.py --batch 64 --frames 243 --lite > gpurun_out/train_lite.json 2> gpurun_out/train_lite.err; cat gpurun_out/train_lite.json; tail -3 gpurun_out/train_lite.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
    python scripts/bench_train.py --batch 16 --frames 243 --steps 1 --warmup 1 --skip-torch > gpurun_out/train_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open("gpurun_out/train_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
n = len(rows) - 1
half = rows[1 + n // 2:]            # second step (after warm-up)
agg = collections.OrderedDict()
for r in half:
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", "")) * (1e-3 if r[ui] in ("ns", "nsecond") else 1.0)   # -> us
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
print(f"step total {tot/1e3:.2f} ms over {sum(c for c,_ in agg.values())} launches")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v/1e3:9.3f} ms {c:5d}x  {k[:110]}")
PY
