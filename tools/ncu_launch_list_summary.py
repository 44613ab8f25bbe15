"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (file argument): kernels by total time."""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
n = 0
for r in rd:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1e-6)
    name = r[ki]
    m = re.search(r"(lyco::\w+)", name)
    short = ("void " + m.group(1)) if m else re.sub(r"<.*", "", name)
    short = short.replace("native::", "").replace("(anonymous namespace)::", "").replace("cudnn::", "")[:78]
    agg[short][0] += 1
    agg[short][1] += v
    total += v
    n += 1
print(f"{n} launches, {total:.2f} ms serialised")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t:9.3f} ms {100 * t / total:5.1f}% {c:6d}  {k}")
