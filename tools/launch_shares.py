#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total ms, share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[r[ki][:90]][0] += 1
    agg[r[ki][:90]][1] += v
tot = sum(v for _, v in agg.values())
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print("%6d %10.3f ms %5.1f%%  %s" % (n, v / 1e6, 100 * v / tot, k))
