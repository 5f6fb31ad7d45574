#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total ms and share."""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    name = r["Kernel Name"]
    name = name[:name.index("(")] if "(" in name else name
    agg[name[:90]][0] += 1
    agg[name[:90]][1] += ms
tot = sum(v[1] for v in agg.values())
print("total device time of listed launches: %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 15]:
    print("%10.3f ms %5dx %6.2f%%  avg %8.3f ms  %s" % (v[1], v[0], 100 * v[1] / tot, v[1] / v[0], k))
