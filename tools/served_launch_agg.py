#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, mean, max, share.  usage: served_launch_agg.py <csv>"""
import csv
import collections
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[ki][:64]].append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
print('%d launches, %.1f ms in total' % (sum(len(v) for v in agg.values()), tot / 1e6))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-66s n=%5d mean %8.1f us  max %8.1f us  share %.3f' % (k, len(v), sum(v) / len(v) / 1e3, max(v) / 1e3, sum(v) / tot))
