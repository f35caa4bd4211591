"""Print the top stall locations (SASS) of one kernel from an .ncu-rep:  ncu_top_stalls.py rep regex [skip] [n]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
k = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    try:
        v = float(r[k])
    except Exception:
        continue
    data.append((v, idx, r))
tot = sum(v for v, _, _ in data)
print("total samples", tot)
for v, idx, r in sorted(data, key=lambda t: -t[0])[:topn]:
    st = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{v:7.0f} {v / tot:6.3f} #{idx:5d} {r[1].strip()[:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
