#!/usr/bin/env python
"""Wall-clock of the host-buffer path (vvcb_frame_begin + vvcb_reco_update + vvcb_rmd_eval, pinned buffers) for one 1080p sweep,
and of its parts."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import vvc_intra_b200 as vb
from make_golden import synth_yuv
W, H = 1920, 1080
Y = synth_yuv(W, H, 10)[0].astype(np.int16)
vis = vb.build_sweep_visits(W, H, qp=32)
with vb.IntraCostEngine(0, 10, 128) as eng:
    hv = eng.host_array(len(vis), vb.VISIT_DTYPE); hv[:] = vis
    hr = eng.host_array(len(vis), vb.RESULT_DTYPE)
    hy = eng.host_array(H * W, np.int16).reshape(H, W); hy[:] = Y
    def timed(fn, reps=5):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        return (time.perf_counter() - t0) / reps * 1e3
    planes = timed(lambda: (eng.frame_begin(hy), eng.reco_update(hy)))
    evalonly = timed(lambda: eng.rmd_eval(hv, out=hr))
    full = timed(lambda: (eng.frame_begin(hy), eng.reco_update(hy), eng.rmd_eval(hv, out=hr)))
    print('chunk', os.environ.get('VVCB_PIPE_CHUNK', 'default'), 'planes %.2f ms, rmd_eval %.2f ms, full step %.2f ms = %.0f CTU/s' % (planes, evalonly, full, 135 / (full * 1e-3)))
