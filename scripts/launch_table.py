#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    k, v, u = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= v:
            continue
        t = float(r[v].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[u], 1.0)
        a = agg.setdefault(r[k][:70], [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f"# {path}")
    for n, (c, t) in agg.items():
        print(f"{c:4d} x {t / c:10.1f} us = {t / 1e3:9.3f} ms  {n}")
