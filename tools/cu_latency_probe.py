#!/usr/bin/env python
"""Round-trip latency of vvcb_cu_eval as a host walk sees it (one CU per request, several requests per call as the broker merges them):
prints mean / p50 / p90 microseconds per call for the rough mode decision alone, the TU candidates alone and both, at a few CU sizes and
batch widths.  GPU box only; a measuring aid for profiles/, not a bench line."""
import ctypes as C
import json
import sys
import time
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tools'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import vvc_intra_b200 as vb
from vvc_intra_b200 import engine as E


def build(eng, rng, lw, lh, n_req, n_jobs, want_rmd, plane=(1024, 2048)):
    keep, arr = [], (E.CuRequest * n_req)()
    w, h = 1 << lw, 1 << lh
    for r in arr:
        v = np.zeros(1, vb.VISIT_DTYPE)
        v['x'] = 4 * rng.integers(2, (plane[1] - 3 * w) // 4)
        v['y'] = 4 * rng.integers(2, (plane[0] - 3 * h) // 4) | 4
        v['log2w'], v['log2h'] = lw, lh
        v['avail_al'], v['n_above'], v['n_above_right'], v['n_left'], v['n_below_left'] = 1, w // 4, w // 4, h // 4, 0
        v['mpm'] = [0, 50, 18, 46, 54, 1]
        v['num_mpm_cand'] = 2
        v['rates'] = rng.integers(1000, 90000, 11)
        v['sqrt_lambda'] = 0.002
        x, y = int(v['x'][0]), int(v['y'][0])
        rects = np.array([(x - 4, y - 4, 2 * w + 8, 4, 0), (x - 4, y, 4, 2 * h + 4, 4 * (2 * w + 8))], vb.RECT_DTYPE)
        smp = rng.integers(0, 1023, 4 * (2 * w + 8) + 4 * (2 * h + 4)).astype(np.int16)
        res, det = np.zeros(1, vb.RESULT_DTYPE), np.zeros(1, vb.DETAIL_DTYPE)
        keep += [v, rects, smp, res, det]
        r.rects, r.n_rects, r.rect_samples, r.n_rect_samples = rects.ctypes.data, 2, smp.ctypes.data, smp.size
        r.visit = v.ctypes.data
        if want_rmd:
            r.want_rmd, r.result, r.detail = 1, res.ctypes.data, det.ctypes.data
        if n_jobs:
            jobs = np.zeros(n_jobs, vb.TU_JOB_DTYPE)
            jobs['x'], jobs['y'], jobs['log2w'], jobs['log2h'] = x, y, lw, lh
            ts = (np.arange(n_jobs) % 5 == 1) & (lw <= 5) & (lh <= 5)
            jobs['mts_idx'] = np.where(ts, 1, 0)
            jobs['lfnst_idx'] = np.where(ts, 0, np.arange(n_jobs) % 3)
            jobs['intra_mode'] = 34
            jobs['flags'] = vb.TU_QUANT | vb.TU_RATE | np.where(ts, vb.TU_RDOQ_TS, vb.TU_DEPQUANT)
            jobs['qp_per'], jobs['qp_rem'] = 7, 2
            jobs['offset'] = np.arange(n_jobs) * w * h
            jobs['lambda'] = 60.0
            slots = (np.arange(n_jobs) * 7 % 67).astype(np.uint8)
            rates, states = vb.default_dq_rates(), vb.default_ctx_states()
            lvl, rec, prd = np.zeros(n_jobs * w * h, np.int32), np.zeros(n_jobs * w * h, np.int16), np.zeros(n_jobs * w * h, np.int16)
            tr = np.zeros(n_jobs, vb.TU_RESULT_DTYPE)
            keep += [jobs, slots, rates, states, lvl, rec, prd, tr]
            r.jobs, r.slots, r.n_jobs, r.rates, r.states = jobs.ctypes.data, slots.ctypes.data, n_jobs, rates.ctypes.data, states.ctypes.data
            r.level, r.reco, r.pred, r.tu_results = lvl.ctypes.data, rec.ctypes.data, prd.ctypes.data, tr.ctypes.data
    return arr, keep


def main():
    rng = np.random.default_rng(5)
    out = []
    with vb.IntraCostEngine(device=0, bit_depth=10, ctu_size=128) as eng:
        plane = (1024, 2048)
        eng.frame_begin(rng.integers(0, 1023, plane).astype(np.int16))
        lib, ctx = eng._lib, eng._ctx
        for (lw, lh) in ((2, 2), (3, 3), (4, 4), (5, 5), (6, 6)):
            for n_req in (1, 8, 32):
                for (want_rmd, n_jobs, name) in ((1, 0, 'rmd'), (0, 32, 'tu32'), (1, 32, 'rmd+tu32')):
                    arr, keep = build(eng, rng, lw, lh, n_req, n_jobs, want_rmd, plane)
                    ts = []
                    for it in range(60):
                        t0 = time.perf_counter()
                        rc = lib.vvcb_cu_eval(ctx, C.byref(arr), n_req)
                        ts.append(time.perf_counter() - t0)
                        if rc:
                            raise SystemExit(lib.vvcb_last_error(ctx).decode())
                    t = np.array(ts[10:]) * 1e6
                    kt = None
                    if n_jobs and os.environ.get('PROBE_KERNELS'):
                        eng.kernel_timing(True)
                        for it in range(10):
                            lib.vvcb_cu_eval(ctx, C.byref(arr), n_req)
                        k = eng.tu_kernel_times()
                        eng.kernel_times()
                        eng.kernel_timing(False)
                        kt = [1e3 * x / max(1, k[4]) for x in k[:4]]
                    out.append(dict(size='%dx%d' % (1 << lw, 1 << lh), requests=n_req, kind=name, mean_us=float(t.mean()), p50_us=float(np.median(t)), p90_us=float(np.percentile(t, 90)), tu_kernel_us=kt))
                    print(out[-1], flush=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'cu_latency.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
