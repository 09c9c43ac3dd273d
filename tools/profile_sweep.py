#!/usr/bin/env python
"""Small driver for ncu: one exhaustive RMD sweep of a WxH synthetic 10-bit frame (after one warm-up pass)."""
import argparse
import os
import sys

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
ap.add_argument('--host', action='store_true', help='go through the host-buffer API (chunked copy/compute pipeline) instead of resident buffers')
a = ap.parse_args()
Y = synth_yuv(a.width, a.height, 10)[0].astype(np.int16)
vis = vb.build_sweep_visits(a.width, a.height, qp=32)
with vb.IntraCostEngine(0, 10, 128) as eng:
    eng.frame_begin(Y)
    eng.reco_update(Y)
    eng.kernel_timing(True)
    if a.host:
        for _ in range(a.passes):
            res = eng.rmd_eval(vis)
    else:                                  # resident: one launch per bucket for the whole frame (what bench.py times as `value`)
        d_vis = eng.dev_alloc(vis.nbytes)
        eng.dev_upload(d_vis, vis)
        d_res = eng.dev_alloc(len(vis) * vb.RESULT_DTYPE.itemsize)
        for _ in range(a.passes):
            eng.rmd_eval_device(d_vis, len(vis), d_res, None)
        eng.sync()
        res = np.zeros(len(vis), vb.RESULT_DTYPE)
        eng.dev_download(res, d_res)
    print('visits', len(vis), 'kernel ms (plan, eval, lists, n):', eng.kernel_times(), 'n_rd[0]', res['n_rd'][0])
