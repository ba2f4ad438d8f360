#!/usr/bin/env python
"""Average duration per kernel of an ncu launch list (--metrics gpu__time_duration.sum --csv).  Usage: tools/launch_summary.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ix = {h: j for j, h in enumerate(hdr)}
d = collections.OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    scale = 1e-3 if r[ix["Metric Unit"]] in ("ns", "nsecond") else 1.0
    d.setdefault(r[ix["Kernel Name"]], []).append(float(r[ix["Metric Value"]].replace(",", "")) * scale)
tot = 0.0
for k, v in d.items():
    print("%-60s n=%4d  avg %8.1f us  min %8.1f  max %8.1f" % (k[:60], len(v), sum(v) / len(v), min(v), max(v)))
