#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean and share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi and r[0].isdigit():
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v
        agg.setdefault(r[ki][:70], []).append(v)
total = sum(sum(v) for v in agg.values())
print(f"{'kernel':72s} {'n':>4s} {'mean us':>10s} {'share':>7s}")
for k, v in agg.items():
    print(f"{k:72s} {len(v):4d} {sum(v)/len(v):10.1f} {100*sum(v)/total:6.1f}%")
