#!/usr/bin/env python
"""Kernel time of one resident 1080p sweep cut into n calls of vvcb_rmd_eval_device (what the chunked host-buffer pipeline pays per
chunk beyond the copies): prints ms per sweep for n = 1, 2, 4, 7, 14."""
import os, sys, time
import numpy as np
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import vvc_intra_b200 as vb
from make_golden import synth_yuv
W,H=1920,1080
Y = synth_yuv(W, H, 10)[0].astype(np.int16)
vis = vb.build_sweep_visits(W, H, qp=32)
n=len(vis)
with vb.IntraCostEngine(0, 10, 128) as eng:
    eng.frame_begin(Y); eng.reco_update(Y)
    d_vis = eng.dev_alloc(vis.nbytes); eng.dev_upload(d_vis, vis)
    d_res = eng.dev_alloc(n * vb.RESULT_DTYPE.itemsize)
    for chunks in (1, 2, 4, 7, 14):
        m = (n + chunks - 1)//chunks
        def run():
            for c in range(chunks):
                off = c*m; k = min(m, n-off)
                eng.rmd_eval_device(d_vis.value + off*80 if hasattr(d_vis,'value') else d_vis + off*80, k, (d_res.value if hasattr(d_res,'value') else d_res) + off*368, None)
        run(); eng.sync()
        eng.timer_start()
        for _ in range(3): run()
        ms = eng.timer_stop()/3
        print('chunks', chunks, 'ms per sweep %.2f' % ms)
