#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel count, last and mean duration."""
import collections, csv, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1], errors="ignore") if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0][:60]
    agg.setdefault(k, []).append(float(r["Metric Value"]) / 1e6)
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} last={v[-1]:9.3f} ms  mean={sum(v)/len(v):9.3f}")
