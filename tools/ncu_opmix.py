#!/usr/bin/env python
"""Instruction mix of one kernel launch from an .ncu-rep source page: executed warp instructions and stall samples per opcode."""
import csv, subprocess, sys, collections
rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:rmd_eval_kernel', '--launch-skip', skip,
                      '--launch-count', '1'], stdout=subprocess.PIPE, text=True).stdout.splitlines()
print(out[0][:120])
rows = list(csv.reader(out[1:]))
hdr = rows[0]; c = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); smp = collections.Counter(); thr = collections.Counter()
for r in rows[1:]:
    if len(r) <= c["Thread Instructions Executed"] or not r[c["Instructions Executed"]].isdigit() or not r[c["Address"]].startswith("0x"): continue
    src = r[c['Source']].strip()
    if src.startswith('@'): src = src.split(None, 1)[1]
    op = src.split()[0] if src else '?'
    op = op if len(sys.argv) > 3 else op.split('.')[0]
    n = int(r[c['Instructions Executed']]); ops[op] += n; smp[op] += int(r[c['# Samples']]); thr[op] += int(r[c['Thread Instructions Executed']])
tot = sum(ops.values()); ts = sum(smp.values())
print('total warp inst %d  avg active threads %.1f  samples %d  static instrs %d' % (tot, sum(thr.values()) / tot, ts, len(rows) - 1))
for op, n in ops.most_common(28):
    print('%-10s %6.2f%% inst   %6.2f%% samples   thr/inst %.1f' % (op, 100.0 * n / tot, 100.0 * smp[op] / max(ts, 1), thr[op] / max(n, 1)))
