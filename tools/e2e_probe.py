#!/usr/bin/env python
"""Wall-clock of the host-buffer path (vvcb_frame_begin + vvcb_reco_update + vvcb_rmd_eval, pinned buffers) for one 1080p sweep."""
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
    def step():
        eng.frame_begin(hy); eng.reco_update(hy); eng.rmd_eval(hv, out=hr)
    step(); step()
    t0 = time.perf_counter()
    for _ in range(5): step()
    dt = (time.perf_counter() - t0) / 5
    print('chunk', os.environ.get('VVCB_PIPE_CHUNK', 'default'), 'e2e ms per sweep %.2f' % (dt * 1e3), 'CTU/s %.0f' % (135 / dt))
