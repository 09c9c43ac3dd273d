#!/usr/bin/env python
"""Joins the SASS-level counters of one kernel launch in an .ncu-rep with nvdisasm line info:
executed warp instructions and stall samples per source line (top N).
usage: ncu_lines.py <rep> <launch-skip among the matching kernels> <mangled substring e.g. ILi3ELi0> [N] [kernel name regex]"""
import csv, subprocess, sys, collections, re, os, glob, tempfile
rep, skip, sub = sys.argv[1], sys.argv[2], sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(root, 'vvc_intra_b200/libvvc_intra_b200.so')], cwd=tmp, stdout=subprocess.DEVNULL)
dis = subprocess.run(['nvdisasm', '--print-line-info', glob.glob(tmp + '/*.cubin')[0]], stdout=subprocess.PIPE, text=True).stdout.splitlines()
line_of, cur, infn = {}, None, False
for l in dis:
    if l.startswith('//--------------------- .text.'):
        infn = sub in l
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + (sys.argv[5] if len(sys.argv) > 5 else 'rmd_eval_kernel'), '--launch-skip', skip,
                      '--launch-count', '1'], stdout=subprocess.PIPE, text=True).stdout.splitlines()
print(out[0][:100])
rows = list(csv.reader(out[1:])); hdr = rows[0]; c = {h: i for i, h in enumerate(hdr)}
base = None
inst = collections.Counter(); smp = collections.Counter(); tot = ts = 0
for r in rows[1:]:
    if len(r) <= c['Thread Instructions Executed'] or not r[c['Address']].startswith('0x') or not r[c['Instructions Executed']].isdigit(): continue
    a = int(r[c['Address']], 16)
    if base is None: base = a
    key = line_of.get(a - base, ('?', 0))
    n = int(r[c['Instructions Executed']]); s = int(r[c['# Samples']])
    inst[key] += n; smp[key] += s; tot += n; ts += s
src = {}
for f in set(k[0] for k in inst if k):
    p = os.path.join(root, 'vvc_intra_b200/csrc', f)
    if os.path.exists(p): src[f] = open(p).read().splitlines()
print('total warp inst', tot, 'samples', ts)
order = smp.most_common(N) if os.environ.get('BY_SAMPLES') else inst.most_common(N)
for key, _ in order:
    n = inst[key]
    f, ln = key if key else ('?', 0)
    text = src.get(f, [''] * (ln + 1))[ln - 1].strip()[:90] if ln else ''
    print('%5.2f%% inst %5.2f%% smp  %s:%d  %s' % (100.0 * n / tot, 100.0 * smp[key] / max(ts, 1), f, ln, text))
