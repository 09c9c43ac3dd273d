#!/usr/bin/env python
"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of the rmd_eval_kernel launches of one
1080p sweep) -> profiles/eval_traffic.json, which bench.py reports as roofline.traffic.
usage: make_traffic.py <csv> <tag>"""
import collections
import csv
import json
import os
import sys

src, tag = sys.argv[1], sys.argv[2]
per = collections.OrderedDict()
unit = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9}
for r in csv.reader(open(src)):
    if len(r) > 14 and r[0].isdigit():
        k = per.setdefault(r[0], dict(kernel=r[4].replace('<unnamed>::', '').split('(')[0].replace('void ', '')))
        k[r[12]] = float(r[14].replace(',', '')) * unit.get(r[13], 1)
rows = list(per.values())
rd = sum(k.get('dram__bytes_read.sum', 0) for k in rows)
wr = sum(k.get('dram__bytes_write.sum', 0) for k in rows)
out = dict(dram_bytes_per_step=rd + wr, dram_read_bytes=rd, dram_write_bytes=wr, launches=len(rows),
           kernel_time_under_ncu_s=sum(k.get('gpu__time_duration.sum', 0) for k in rows),
           source='ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, the %d rmd_eval_kernel launches of one '
                  '1920x1080 sweep (tools/profile_sweep.py, second pass), round %s' % (len(rows), tag),
           per_launch=rows)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, 'gpurun_out', 'eval_traffic_%s.json' % tag), 'w'), indent=1)
print('dram read %.1f MB write %.1f MB over %d launches' % (rd / 1e6, wr / 1e6, len(rows)))
