"""Summarise an `ncu --metrics gpu__time_duration.sum` launch list of scripts/bench_train.py (warm-up step + one
measured step): per-kernel totals of the LAST step, written as JSON for profiles/."""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
body = rows[1:]
# the step boundary is the embed kernel of the second forward
starts = [i for i, r in enumerate(body) if "embed_kernel" in r[ki]]
step = body[starts[-1]:]
agg = collections.OrderedDict()
for r in step:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mb::", "")
    v = float(r[vi].replace(",", "")) * (1e-3 if r[ui].startswith("n") else 1.0)      # -> us
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
out = {"source": src, "launches": sum(c for c, _ in agg.values()), "total_ms_cold_serialised": tot / 1e3,
       "kernels": [{"kernel": k, "launches": c, "ms": v / 1e3, "share": v / tot}
                   for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open(dst, "w"), indent=1)
print(f"step: {out['launches']} launches, {tot / 1e3:.2f} ms (cold, serialised)")
for k in out["kernels"][:30]:
    print(f"{k['ms']:9.3f} ms {k['launches']:5d}x {k['share'] * 100:5.1f}%  {k['kernel'][:100]}")
