"""Per-kernel histogram of the Blackwell-specific SASS mnemonics in the shipped library (B200_PROFILING.md table):
   python scripts/sass_histogram.py [lib.so] > profiles/rNN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "motionbert_b200", "libmotionbert_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", sass)
if names:
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2",
        "F2FP", "RED", "UTCATOMSWS"]
cur, per = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for w in WANT:
            if op.startswith(w):
                key = op if w in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTCBAR") else w
                per[cur][key] += 1
tot = collections.Counter()
print(f"# SASS mnemonic histogram of {os.path.basename(lib)} ({len(per)} kernels); UTC*MMA = tcgen05.mma (UTCHMMA: kind::f16, UTCQMMA:")
print("# kind::f8f6f4), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, HMMA = legacy mma.sync (must be 0)")
for k, c in per.items():
    name = re.sub(r"\(CUtensorMap.*|\(const.*|\(mb::.*", "", demangle.get(k, k)).replace("void mb::", "")
    items = ", ".join(f"{a} {b}" for a, b in sorted(c.items()) if a != "_total")
    print(f"{name:70s} instrs {c['_total']:6d}  {items}")
    tot.update({a: b for a, b in c.items()})
print("# TOTAL " + ", ".join(f"{a} {b}" for a, b in sorted(tot.items())))
