"""Hot spots of one kernel from an exported `ncu --page source --csv` file:  ncu_source_csv.py file.csv kernel_regex [nth] [top]"""
import csv
import re
import sys

f, rx = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(f)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = [i for i in starts if re.search(rx, rows[i][1])]
s = sel[nth]
e = min([i for i in starts if i > s] + [len(rows)])
hdr = rows[s + 1]
ks, ki = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[s + 2:e]:
    try:
        data.append((float(r[ks] or 0), float(r[ki] or 0), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
toti = sum(d[1] for d in data)
print(rows[s][1][:90], "| samples", tot, "| warp instructions", toti)
agg = {}
for v, n, r in data:
    for i in stall:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + float(r[i] or 0)
print("stall mix:", ", ".join(f"{k} {v / max(tot, 1):.2f}" for k, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
ops = {}
for v, n, r in data:
    op = r[1].strip().split()[0] if r[1].strip() else "?"
    if op.startswith("@"):
        op = r[1].strip().split()[1]
    op = op.split(".")[0]
    o = ops.setdefault(op, [0, 0])
    o[0] += n
    o[1] += v
print("opcode mix (warp instrs, samples):", ", ".join(f"{k} {a / max(toti, 1):.3f}/{b / max(tot, 1):.3f}" for k, (a, b) in sorted(ops.items(), key=lambda t: -t[1][0])[:18]))
for v, n, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{v:7.0f} {v / max(tot, 1):6.3f} {n:9.0f} {r[1].strip()[:76]:76s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
