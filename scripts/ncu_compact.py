"""One line per profiled launch from `ncu --page raw --csv` dumps:  ncu_compact.py a_raw.csv [b_raw.csv ...] [--json out]"""
import csv
import json
import re
import sys

W = {"gpu__time_duration.sum": "us", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
     "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
     "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex%",
     "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr", "launch__registers_per_thread": "regs",
     "launch__grid_size": "grid", "launch__block_size": "block", "sm__cycles_elapsed.avg.per_second": "ghz"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
args = [a for a in sys.argv[1:] if a != "--json"]
jout = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
if jout:
    args.remove(jout)
allk = []
for f in args:
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        print(f, "empty")
        continue
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        name = re.sub(r"\(CUtensorMap.*|\(const.*|\(mb::.*", "", r[h.index("Kernel Name")]).replace("void mb::", "")
        k = {"kernel": name, "file": f.split("/")[-1]}
        for i, col in enumerate(h):
            if col in W:
                v = float(r[i].replace(",", ""))
                if W[col] in ("rd", "wr", "us"):
                    v *= UNIT.get(u[i], 1.0)
                k[W[col]] = v
        k["GB"] = (k.get("rd", 0) + k.get("wr", 0)) / 1e9
        k["TBps"] = k["GB"] / k["us"] / 1e3 * 1e3 / 1e3 if k.get("us") else 0
        allk.append(k)
        print(f"{name[:46]:46s} {k.get('us', 0):9.1f} us  tensor {k.get('tensor%', 0):5.1f}%  issue {k.get('issue%', 0):5.1f}%  "
              f"dram {k.get('dram%', 0):5.1f}% ({k['GB']:.2f} GB)  l1tex {k.get('l1tex%', 0):5.1f}%  regs {int(k.get('regs', 0))}  "
              f"grid {int(k.get('grid', 0))}x{int(k.get('block', 0))}")
if jout:
    json.dump(allk, open(jout, "w"), indent=1)
