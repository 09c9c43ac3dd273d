#!/usr/bin/env python
"""Small driver for timing / ncu: the TU-coding sweep (forward transform, dependent quantisation, reconstruction, SSE) of the
candidate CUs of a WxH synthetic 10-bit frame."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import vvc_intra_b200 as vb          # noqa: E402
from make_golden import synth_yuv    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--width', type=int, default=416)
ap.add_argument('--height', type=int, default=240)
ap.add_argument('--passes', type=int, default=2)
ap.add_argument('--qp', type=int, default=32)
ap.add_argument('--scalar', action='store_true')
a = ap.parse_args()
Y = synth_yuv(a.width, a.height, 10)[0].astype(np.int16)
vis = vb.build_sweep_visits(a.width, a.height, qp=a.qp)
t0 = time.perf_counter()
jobs, resi, pred, rates = vb.build_tu_sweep(Y, vis, a.qp, 10, dep_quant=not a.scalar)
print('jobs', len(jobs), 'samples', resi.size, 'build %.2fs' % (time.perf_counter() - t0))
with vb.IntraCostEngine(0, 10, 128) as eng:
    eng.frame_begin(Y)
    eng.kernel_timing(True)
    for _ in range(a.passes):
        t0 = time.perf_counter()
        out = eng.tu_eval(jobs, resi, pred, rates=rates)
        dt = time.perf_counter() - t0
    r = out['results']
    print('kernel ms (transform, quantisers, recon, rate, n):', eng.tu_kernel_times(), 'last call %.1f ms' % (dt * 1e3),
          'nonzero TUs %.3f' % float((r['abs_sum_level'] > 0).mean()), 'mean abs sum %.2f' % float(r['abs_sum_level'].mean()))
