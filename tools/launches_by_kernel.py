"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launches_by_kernel.py launches.csv 'header line'"""
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    us = float(r[v].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[u], 1e-3)
    name = re.sub(r"\(.*", "", r[k])[:90]
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + us)
tot = sum(t for _, t in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("%-92s %6s %12s  %s" % ("kernel", "count", "total us", "share"))
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-92s %6d %12.1f  %4.1f%%" % (name, c, t, 100 * t / tot))
