"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py launches.csv [nsteps]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    name = r[ik].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in agg.values())
print("# kernel, launches (%d steps), total_us (%d steps), share   [sum %.1f us = %.3f ms/step]"
      % (nsteps, nsteps, tot, tot / nsteps / 1e3))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s, %d, %.1f, %.3f" % (k, n, v, v / tot))
