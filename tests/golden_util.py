"""Helpers shared by the oracle and GPU parity tests: load the committed reference traces and lay the
recorded visits out on one synthetic "atlas" picture so that a whole fixture is one batch."""
import functools
import gzip
import os

import numpy as np

from trace_parse import iter_records, group_visits
from oracle import oracle_py as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
KFAST = [[3, 3, 3, 3, 2, 2], [3, 3, 3, 3, 3, 2], [3, 3, 3, 3, 3, 2], [3, 3, 3, 3, 3, 2], [2, 3, 3, 3, 3, 2], [2, 2, 2, 2, 2, 3]]


@functools.lru_cache(maxsize=None)
def load_fixture(name):
    raw = gzip.open(os.path.join(GOLDEN, name + '.bin.gz'), 'rb').read()
    return group_visits(iter_records(raw))


def slot_of(head, e):
    if e['mip']:
        return O.SLOT_MIP + e['mode']
    if e['mrl']:
        return (O.SLOT_MRL1 if e['mrl'] == 1 else O.SLOT_MRL3) + list(head['mpm'][1:]).index(e['mode'])
    return e['mode']


def build_atlas(visits, atlas_w=2048):
    """Places every recorded visit's reconstruction window and original block in one picture.
    Returns (orig, reco, visit_array).  The cell of a visit is (2w+8) x (2h+8); the CU sits at +4,+4."""
    cells, x, y, rowh = [], 0, 0, 0
    for v in visits:
        w, h = v['head']['w'], v['head']['h']
        cw, ch = 2 * w + 8, 2 * h + 8
        if x + cw > atlas_w:
            x, y, rowh = 0, y + rowh, 0
        cells.append((x, y))
        x += cw
        rowh = max(rowh, ch)
    H = y + rowh
    orig = np.zeros((H, atlas_w), np.int16)
    reco = np.zeros((H, atlas_w), np.int16)
    arr = np.zeros(len(visits), O.VISIT_DTYPE)
    for i, (v, (cx, cy)) in enumerate(zip(visits, cells)):
        hd, r = v['head'], v['refs'][0]
        w, h = hd['w'], hd['h']
        reco[cy:cy + 4, cx:cx + 2 * w + 8] = r['reco_top']
        reco[cy + 4:cy + 2 * h + 8, cx:cx + 4] = r['reco_left']
        orig[cy + 4:cy + 4 + h, cx + 4:cx + 4 + w] = hd['org']
        a = arr[i]
        a['x'], a['y'] = cx + 4, cy + 4
        a['log2w'], a['log2h'] = w.bit_length() - 1, h.bit_length() - 1
        for k in ('avail_al', 'n_above', 'n_above_right', 'n_left', 'n_below_left'):
            a[k] = r[k]
        # the atlas position does not keep y & (ctu-1): carry "first line of CTU" as a flag instead
        a['flags'] = 1 if (hd['y'] & 127) == 0 else 0
        a['mpm'] = hd['mpm']
        a['num_mpm_cand'] = hd['num_cand_mpm']
        a['rates'] = hd['rates']
        a['sqrt_lambda'] = hd['sqrt_lambda']
    return orig, reco, arr


def result_list(res, pre, cost=True):
    n = int(res['n_' + pre])
    out = []
    for i in range(n):
        m = res[pre + '_mode'][i]
        t = (int(m['mip']), int(m['mrl']), int(m['mode']))
        out.append(t + ((float(res[pre + '_cost'][i]),) if cost else ()))
    return out


def check_visit_against_reference(v, res, det):
    """Compares one RMD result (oracle or GPU) with what the reference encoder produced.  Returns a
    list of human-readable mismatches (empty = parity)."""
    hd = v['head']
    w, h = hd['w'], hd['h']
    errs = []
    for e in v['evals']:
        s = slot_of(hd, e)
        if int(det['sad'][s]) != e['sad'] or int(det['satd'][s]) != e['satd']:
            errs.append('slot %d %dx%d sad %d/%d satd %d/%d' % (s, w, h, det['sad'][s], e['sad'], det['satd'][s], e['satd']))
    L = v['lists']
    exp_rd = [(a['mip'], a['mrl'], a['mode'], a['cost']) for a in L['rd']]
    exp_had = [(a['mip'], a['mrl'], a['mode'], a['cost']) for a in L['had']]
    if L['variant'] == 0:
        got_rd, got_had = result_list(res, 'rd'), result_list(res, 'had')
    else:   # EL/IntraSearch.cpp:686-701 saved the regular-only lists, truncated
        k = KFAST[w.bit_length() - 3][h.bit_length() - 3]
        got_rd, got_had = result_list(det, 'reg')[:k], result_list(det, 'reg_had')[:3]
    if got_rd != exp_rd:
        errs.append('rd list %dx%d variant %d: %r != %r' % (w, h, L['variant'], got_rd, exp_rd))
    if got_had != exp_had:
        errs.append('had list %dx%d variant %d: %r != %r' % (w, h, L['variant'], got_had, exp_had))
    if L['final']:
        exp = [(a['mip'], a['mrl'], a['mode']) for a in L['final']]
        got = result_list(res, 'final', cost=False)
        if got != exp:
            errs.append('final list %dx%d: %r != %r' % (w, h, got, exp))
    return errs


def random_case(rng, bd, n_per_shape, plane=(512, 1024)):
    H, W = plane
    orig = rng.integers(0, 1 << bd, (H, W)).astype(np.int16)
    reco = rng.integers(0, 1 << bd, (H, W)).astype(np.int16)
    vis = []
    for lw in range(2, 7):
        for lh in range(2, 7):
            if (lw == 6) != (lh == 6):
                continue
            w, h = 1 << lw, 1 << lh
            for _ in range(n_per_shape):
                v = np.zeros(1, O.VISIT_DTYPE)[0]
                v['x'] = 4 * rng.integers(1, (W - 2 * w) // 4)
                v['y'] = 4 * rng.integers(1, (H - 2 * h) // 4)
                v['log2w'], v['log2h'] = lw, lh
                mode = rng.integers(0, 4)
                if mode == 0:        # everything available
                    v['avail_al'], v['n_above'], v['n_above_right'], v['n_left'], v['n_below_left'] = 1, w // 4, w // 4, h // 4, h // 4
                elif mode == 1:      # nothing
                    pass
                else:                # ragged
                    v['avail_al'] = rng.integers(0, 2)
                    v['n_above'] = rng.integers(0, w // 4 + 1)
                    v['n_above_right'] = rng.integers(0, w // 4 + 1)
                    v['n_left'] = rng.integers(0, h // 4 + 1)
                    v['n_below_left'] = rng.integers(0, h // 4 + 1)
                v['flags'] = rng.integers(0, 4) if rng.random() < 0.3 else 0
                L, A = rng.integers(0, 67, 2)
                mpm, nc = O.intra_mpms(int(L), int(A))
                v['mpm'], v['num_mpm_cand'] = mpm, nc
                v['rates'] = rng.integers(100, 200000, 11)
                v['sqrt_lambda'] = float(rng.uniform(1e-4, 3e-3))
                vis.append(v)
    return orig, reco, np.array(vis, O.VISIT_DTYPE)
