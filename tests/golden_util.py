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


# ---- TU coding: batches built from the reference's 'S' / 'Q' / 'I' records ---------------------------------
def _tu_types():
    import vvc_intra_b200 as vb          # struct layouts only (no library call)
    return vb.TU_JOB_DTYPE, vb.TU_QUANT


def build_tu_batch(tus, bd, seed=0, max_jobs=None):
    """One job per (S record, candidate) [transform + pre-selection sum] and per Q record of a run without dependent
    quantisation [transform + Quant::quant + dequant + inverse + reconstruction + SSE].  The recorded residuals are
    split into a synthetic prediction and original (orig = pred + resi) laid out on an atlas of 64x64 cells.
    Returns (orig_plane, jobs, resi_flat, pred_flat, expect) where expect[i] holds what the reference produced."""
    job_dt, TU_QUANT = _tu_types()
    rng = np.random.default_rng(seed)
    items = []
    for r in tus:
        if r['bd'] != bd:
            continue
        if r['tag'] == 'S':
            for m in r['modes']:
                items.append(dict(kind='S', rec=r, mts=m['mts'], coeff=m['coeff'], resi=r['resi']))
        elif r['tag'] == 'Q' and r['dep_quant'] == 0 and r['lfnst'] == 0:
            items.append(dict(kind='Q', rec=r, mts=r['mts'], coeff=r['coeff'], resi=r['resi'], level=r['level']))
    if max_jobs:
        items = items[:max_jobs]
    n = len(items)
    cols = 16
    orig = np.zeros((64 * ((n + cols - 1) // cols + 1), 64 * cols), np.int16)
    jobs = np.zeros(n, job_dt)
    resi_flat, pred_flat, off = [], [], 0
    mx = (1 << bd) - 1
    for i, it in enumerate(items):
        resi = it['resi']
        h, w = resi.shape
        # prediction chosen so that orig = pred + resi is a legal picture wherever possible (clipped otherwise: the
        # residual handed to the kernel stays the recorded one, as in the reference where resi = org - pred)
        pred = np.clip(rng.integers(0, mx + 1, (h, w)), np.maximum(0, -resi), np.minimum(mx, mx - resi)).astype(np.int16)
        org = np.clip(pred.astype(np.int32) + resi, 0, mx).astype(np.int16)
        cx, cy = 64 * (i % cols), 64 * (i // cols)
        orig[cy:cy + h, cx:cx + w] = org
        j = jobs[i]
        j['x'], j['y'], j['log2w'], j['log2h'] = cx, cy, w.bit_length() - 1, h.bit_length() - 1
        j['mts_idx'] = it['mts']
        j['offset'] = off
        if it['kind'] == 'Q':
            j['flags'] = TU_QUANT
            j['qp_per'], j['qp_rem'] = it['rec']['per'], it['rec']['rem']
        it['pred'], it['org'], it['off'] = pred, org, off
        resi_flat.append(resi.ravel())
        pred_flat.append(pred.ravel())
        off += w * h
    return orig, jobs, np.concatenate(resi_flat), np.concatenate(pred_flat), items


def check_tu_outputs(items, bd, out):
    """out: dict(results, coeff, level, reco) from the kernel.  Compares with the reference records; the dequant + inverse +
    reconstruction + SSE part with the oracle chain (pinned by the 'I' records in tests/test_oracle_tu.py)."""
    errs = []
    for i, it in enumerate(items):
        h, w = it['resi'].shape
        sl = slice(it['off'], it['off'] + w * h)
        r = out['results'][i]
        if not np.array_equal(out['coeff'][sl].reshape(h, w), it['coeff']):
            errs.append('job %d %dx%d mts %d: coefficients differ' % (i, w, h, it['mts']))
        exp_sum = O.abs_sum_for_preselection(it['coeff'], it['mts'])
        if int(r['abs_sum_coeff']) != exp_sum:
            errs.append('job %d: pre-selection sum %d != %d' % (i, r['abs_sum_coeff'], exp_sum))
        if it['kind'] == 'Q':
            rec = it['rec']
            if not np.array_equal(out['level'][sl].reshape(h, w), it['level']):
                errs.append('job %d %dx%d mts %d: levels differ' % (i, w, h, it['mts']))
            if int(r['abs_sum_level']) != rec['abs_sum']:
                errs.append('job %d: abs sum %d != %d' % (i, r['abs_sum_level'], rec['abs_sum']))
            co = O.dequant(it['level'], bd, rec['per'], rec['rem'], it['mts'] == 1)
            res = O.inv_transform(co, bd, it['mts'])
            reco, sse = O.reconstruct_sse(it['org'], it['pred'], res, bd)
            if not np.array_equal(out['reco'][sl].reshape(h, w), reco):
                errs.append('job %d %dx%d mts %d: reconstruction differs' % (i, w, h, it['mts']))
            if int(r['sse']) != int(sse):
                errs.append('job %d: sse %d != %d' % (i, r['sse'], sse))
    return errs


def random_tu_case(rng, bd, n_per_kind, amp=None):
    """Random residuals for every (shape, transform) the path allows (incl. 64-point sides), random QP."""
    job_dt, TU_QUANT = _tu_types()
    mx = (1 << bd) - 1
    items = []
    for lw in range(2, 7):
        for lh in range(2, 7):
            for mts in range(6):
                if mts >= 1 and (lw > 5 or lh > 5):
                    continue
                for _ in range(n_per_kind):
                    w, h = 1 << lw, 1 << lh
                    a = amp if amp is not None else int(rng.choice([3, 30, 300, mx]))
                    a = min(a, mx)
                    pred = rng.integers(0, mx + 1, (h, w))
                    org = np.clip(pred + rng.integers(-a, a + 1, (h, w)), 0, mx)
                    qp = int(rng.integers(0, 64)) + 6 * (bd - 8)
                    items.append(dict(pred=pred.astype(np.int16), org=org.astype(np.int16), resi=(org - pred).astype(np.int16), mts=mts,
                                      per=qp // 6, rem=qp % 6))
    n = len(items)
    cols = 16
    orig = np.zeros((64 * ((n + cols - 1) // cols + 1), 64 * cols), np.int16)
    jobs = np.zeros(n, job_dt)
    off = 0
    for i, it in enumerate(items):
        h, w = it['resi'].shape
        cx, cy = 64 * (i % cols), 64 * (i // cols)
        orig[cy:cy + h, cx:cx + w] = it['org']
        j = jobs[i]
        j['x'], j['y'], j['log2w'], j['log2h'], j['mts_idx'], j['flags'] = cx, cy, w.bit_length() - 1, h.bit_length() - 1, it['mts'], TU_QUANT
        j['qp_per'], j['qp_rem'], j['offset'] = it['per'], it['rem'], off
        it['off'] = off
        off += w * h
    return orig, jobs, np.concatenate([it['resi'].ravel() for it in items]), np.concatenate([it['pred'].ravel() for it in items]), items


def oracle_tu_chain(items, bd):
    """The oracle's whole TU chain for random_tu_case items -> dict like the kernel's outputs."""
    import vvc_intra_b200 as vb
    n = sum(it['resi'].size for it in items)
    out = dict(results=np.zeros(len(items), vb.TU_RESULT_DTYPE), coeff=np.zeros(n, np.int32), level=np.zeros(n, np.int32), reco=np.zeros(n, np.int16))
    for i, it in enumerate(items):
        h, w = it['resi'].shape
        sl = slice(it['off'], it['off'] + w * h)
        co = O.fwd_transform(it['resi'], bd, it['mts'])
        lvl, s = O.quant_scalar(co, bd, it['per'], it['rem'], it['mts'] == 1)
        res = O.inv_transform(O.dequant(lvl, bd, it['per'], it['rem'], it['mts'] == 1), bd, it['mts'])
        reco, sse = O.reconstruct_sse(it['org'], it['pred'], res, bd)
        out['coeff'][sl], out['level'][sl], out['reco'][sl] = co.ravel(), lvl.ravel(), reco.ravel()
        out['results'][i] = (O.abs_sum_for_preselection(co, it['mts']), s, sse, it.get('bits_fn', lambda lv: 0)(lvl))
    return out


# ---- dependent quantisation: batches from the reference's 'D' records ---------------------------------------
def build_dq_batch(tus, bd, seed=0, tag='D'):
    """One VVCB_TU_QUANT | VVCB_TU_DEPQUANT job per 'D' record (or per 'F' record: the same with LFNST); every record brings its
    own context-price snapshot.  Returns (orig_plane, jobs, resi_flat, pred_flat, rates, items)."""
    import vvc_intra_b200 as vb
    rng = np.random.default_rng(seed)
    recs = [r for r in tus if r['tag'] == tag and r['bd'] == bd]
    n = len(recs)
    cols = 16
    orig = np.zeros((64 * ((n + cols - 1) // cols + 1), 64 * cols), np.int16)
    jobs = np.zeros(n, vb.TU_JOB_DTYPE)
    rates = np.zeros(n, vb.DQ_RATES_DTYPE)
    mx = (1 << bd) - 1
    items, off = [], 0
    for i, r in enumerate(recs):
        resi = r['resi']
        h, w = resi.shape
        pred = np.clip(rng.integers(0, mx + 1, (h, w)), np.maximum(0, -resi), np.minimum(mx, mx - resi)).astype(np.int16)
        org = np.clip(pred.astype(np.int32) + resi, 0, mx).astype(np.int16)
        cx, cy = 64 * (i % cols), 64 * (i // cols)
        orig[cy:cy + h, cx:cx + w] = org
        j = jobs[i]
        j['x'], j['y'], j['log2w'], j['log2h'], j['mts_idx'] = cx, cy, w.bit_length() - 1, h.bit_length() - 1, r['mts']
        j['flags'] = vb.TU_QUANT | vb.TU_DEPQUANT
        j['qp_per'], j['qp_rem'], j['offset'] = r['per'], r['rem'], off
        j['rate_idx'], j['lfnst_idx'], j['cbf_delta_bits'], j['lambda'] = i, r['lfnst'], r['cbf_delta'], r['lambda']
        j['intra_mode'] = r.get('intra_mode', 0)
        rates[i] = O.dq_rates_from_flat(r['rates'])
        items.append(dict(rec=r, pred=pred, org=org, off=off))
        off += w * h
    return orig, jobs, np.concatenate([r['resi'].ravel() for r in recs]), np.concatenate([it['pred'].ravel() for it in items]), rates, items


def check_dq_outputs(items, bd, out):
    """Coefficients, levels and absSum against the reference's records; reconstruction + SSE against the oracle's
    state-machine dequantiser followed by the (pinned) inverse transform."""
    errs = []
    for i, it in enumerate(items):
        r = it['rec']
        h, w = r['resi'].shape
        sl = slice(it['off'], it['off'] + w * h)
        res = out['results'][i]
        got_co = out['coeff'][sl].reshape(h, w)
        if r['lfnst']:
            # the reference's buffer holds stale values outside the LFNST region (tests/test_oracle_lfnst.py); ours holds zeros
            sb = 8 if min(w, h) >= 8 else 4
            if not np.array_equal(got_co[:sb, :sb], r['coeff'][:sb, :sb]) or got_co[sb:, :].any() or got_co[:, sb:].any():
                errs.append('job %d %dx%d lfnst %d mode %d: coefficients differ' % (i, w, h, r['lfnst'], r['intra_mode']))
        elif not np.array_equal(got_co, r['coeff']):
            errs.append('job %d %dx%d mts %d: coefficients differ' % (i, w, h, r['mts']))
        if not np.array_equal(out['level'][sl].reshape(h, w), r['level']):
            errs.append('job %d %dx%d mts %d: levels differ (%d positions)' % (i, w, h, r['mts'], int((out['level'][sl].reshape(h, w) != r['level']).sum())))
        if int(res['abs_sum_level']) != r['abs_sum']:
            errs.append('job %d: abs sum %d != %d' % (i, res['abs_sum_level'], r['abs_sum']))
        deq = O.dep_dequant(r['level'], bd, r['qp'])
        if r['lfnst']:
            deq = O.inv_lfnst(deq, r['intra_mode'], r['lfnst'])
        resi = O.inv_transform(deq, bd, r['mts'])
        reco, sse = O.reconstruct_sse(it['org'], it['pred'], resi, bd)
        if not np.array_equal(out['reco'][sl].reshape(h, w), reco):
            errs.append('job %d %dx%d mts %d: reconstruction differs' % (i, w, h, r['mts']))
        if int(res['sse']) != int(sse):
            errs.append('job %d: sse %d != %d' % (i, res['sse'], sse))
    return errs


def random_dq_case(rng, bd, n_per_kind):
    """Random residuals, QP, lambda and context prices for every (shape, transform) dependent quantisation accepts."""
    import vvc_intra_b200 as vb
    mx = (1 << bd) - 1
    items = []
    for lw in range(2, 7):
        for lh in range(2, 7):
            for mts in (0, 2, 3, 4, 5):
                if mts > 1 and (lw > 5 or lh > 5):
                    continue
                for _ in range(n_per_kind):
                    w, h = 1 << lw, 1 << lh
                    a = min(int(rng.choice([2, 12, 80, 400])), mx)
                    pred = rng.integers(0, mx + 1, (h, w))
                    smooth = rng.integers(-a, a + 1, (h // 4 + 1, w // 4 + 1)).repeat(4, 0).repeat(4, 1)[:h, :w]
                    org = np.clip(pred + smooth + rng.integers(-a // 2 - 1, a // 2 + 2, (h, w)), 0, mx)
                    qp = int(rng.integers(10, 52)) + 6 * (bd - 8)
                    items.append(dict(pred=pred.astype(np.int16), org=org.astype(np.int16), resi=(org - pred).astype(np.int16), mts=mts,
                                      qp=qp, lam=float(rng.uniform(0.3, 3.0) * 0.57 * 2.0 ** ((qp - 6 * (bd - 8) - 12) / 3.0)),
                                      cbf=int(rng.integers(-40000, 40000)), lfnst=int(rng.integers(1, 3)) if rng.random() < 0.25 else 0,
                                      intra=int(rng.integers(0, 67))))
    n = len(items)
    cols = 16
    orig = np.zeros((64 * ((n + cols - 1) // cols + 1), 64 * cols), np.int16)
    jobs = np.zeros(n, vb.TU_JOB_DTYPE)
    n_rates = 7
    rates = np.zeros(n_rates, vb.DQ_RATES_DTYPE)
    for name in rates.dtype.names:
        rates[name] = rng.integers(300, 140000, rates[name].shape)
    off = 0
    for i, it in enumerate(items):
        h, w = it['resi'].shape
        cx, cy = 64 * (i % cols), 64 * (i // cols)
        orig[cy:cy + h, cx:cx + w] = it['org']
        j = jobs[i]
        j['x'], j['y'], j['log2w'], j['log2h'], j['mts_idx'] = cx, cy, w.bit_length() - 1, h.bit_length() - 1, it['mts']
        j['flags'] = vb.TU_QUANT | vb.TU_DEPQUANT
        j['qp_per'], j['qp_rem'], j['offset'] = it['qp'] // 6, it['qp'] % 6, off
        j['rate_idx'], j['lfnst_idx'], j['cbf_delta_bits'], j['lambda'] = i % n_rates, it['lfnst'], it['cbf'], it['lam']
        j['intra_mode'] = it['intra']
        it['off'], it['rate'] = off, rates[i % n_rates]
        off += w * h
    return orig, jobs, np.concatenate([it['resi'].ravel() for it in items]), np.concatenate([it['pred'].ravel() for it in items]), rates, items


def oracle_dq_chain(items, bd):
    import vvc_intra_b200 as vb
    n = sum(it['resi'].size for it in items)
    out = dict(results=np.zeros(len(items), vb.TU_RESULT_DTYPE), coeff=np.zeros(n, np.int32), level=np.zeros(n, np.int32), reco=np.zeros(n, np.int16))
    for i, it in enumerate(items):
        h, w = it['resi'].shape
        sl = slice(it['off'], it['off'] + w * h)
        co = O.fwd_transform(it['resi'], bd, it['mts'], it['lfnst'])
        if it['lfnst']:
            co = O.fwd_lfnst(co, it['intra'], it['lfnst'])
        lvl, s = O.dep_quant(co, bd, it['mts'], it['lfnst'], it['qp'], it['lam'], it['rate'], it['cbf'])
        deq = O.dep_dequant(lvl, bd, it['qp'])
        if it['lfnst']:
            deq = O.inv_lfnst(deq, it['intra'], it['lfnst'])
        res = O.inv_transform(deq, bd, it['mts'])
        reco, sse = O.reconstruct_sse(it['org'], it['pred'], res, bd)
        out['coeff'][sl], out['level'][sl], out['reco'][sl] = co.ravel(), lvl.ravel(), reco.ravel()
        out['results'][i] = (O.abs_sum_for_preselection(co, it['mts']), s, sse, it.get('bits_fn', lambda lv: 0)(lvl))
    return out


# ---- TU coding with device-side prediction (vvcb_tu_eval_pred) ---------------------------------------------
def pred_tu_case(rng, bd, n_per_shape, slots_per_visit=4):
    """Random visits (ragged availability) and, per visit, a few evaluation slots of every kind that the visit evaluates,
    each turned into one dependent-quantisation TU job.  Returns (orig, reco, visits, src, jobs, n_samples, rates, expect)
    where expect[i] = (prediction block, oracle TU-chain item)."""
    import vvc_intra_b200 as vb
    orig, reco, visits = random_case(rng, bd, n_per_shape, plane=(256, 512))
    _, _, preds = O.rmd_batch(orig, reco, bd, 128, visits, want_pred=True)
    n_rates = 3
    rates = np.zeros(n_rates, vb.DQ_RATES_DTYPE)
    for name in rates.dtype.names:
        rates[name] = rng.integers(300, 140000, rates[name].shape)
    src, jobs, items, off = [], [], [], 0
    for vi, v in enumerate(visits):
        w, h = 1 << int(v['log2w']), 1 << int(v['log2h'])
        mrl_ok = not (int(v['flags']) & 1) and (int(v['y']) & 127) != 0      # VVCB_VISIT_NO_MRL
        n_mip = 0 if (int(v['flags']) & 2) or w > 4 * h or h > 4 * w else (35 if w == h == 4 else 19 if max(w, h) <= 8 else 11)
        cand = list(rng.choice(67, 2, replace=False)) + [0, 1]
        if mrl_ok:
            cand += [O.SLOT_MRL1 + int(rng.integers(0, 5)), O.SLOT_MRL3 + int(rng.integers(0, 5))]
        if n_mip:
            cand += [O.SLOT_MIP + int(rng.integers(0, n_mip))]
        for slot in rng.permutation(cand)[:slots_per_visit]:
            slot = int(slot)
            mts = int(rng.choice([0, 0, 2, 5])) if max(w, h) <= 32 else 0
            qp = int(rng.integers(16, 48)) + 6 * (bd - 8)
            j = np.zeros(1, vb.TU_JOB_DTYPE)[0]
            j['x'], j['y'], j['log2w'], j['log2h'], j['mts_idx'] = v['x'], v['y'], v['log2w'], v['log2h'], mts
            j['flags'] = vb.TU_QUANT | vb.TU_DEPQUANT
            j['qp_per'], j['qp_rem'], j['offset'] = qp // 6, qp % 6, off
            lam = float(0.57 * 2.0 ** ((qp - 6 * (bd - 8) - 12) / 3.0))
            j['rate_idx'], j['cbf_delta_bits'], j['lambda'] = len(jobs) % n_rates, int(rng.integers(-30000, 30000)), lam
            p = preds[vi][slot]
            o = orig[int(v['y']):int(v['y']) + h, int(v['x']):int(v['x']) + w]
            items.append(dict(pred=p, org=o, resi=(o.astype(np.int32) - p).astype(np.int16), mts=mts, qp=qp, lam=lam, cbf=int(j['cbf_delta_bits']),
                              lfnst=0, intra=0, off=off, rate=rates[len(jobs) % n_rates]))
            s = np.zeros(1, vb.TU_SRC_DTYPE)[0]
            s['visit'], s['slot'] = vi, slot
            src.append(s)
            jobs.append(j)
            off += w * h
    return orig, reco, visits, np.array(src, vb.TU_SRC_DTYPE), np.array(jobs, vb.TU_JOB_DTYPE), off, rates, items


# ---- RDOQ of transform-skip TUs: batches from the reference's 'T' records -----------------------------------
def build_rdoq_batch(tus, bd, seed=0):
    import vvc_intra_b200 as vb
    rng = np.random.default_rng(seed)
    recs = [r for r in tus if r['tag'] == 'T' and r['bd'] == bd]
    n = len(recs)
    cols = 16
    orig = np.zeros((64 * ((n + cols - 1) // cols + 1), 64 * cols), np.int16)
    jobs = np.zeros(n, vb.TU_JOB_DTYPE)
    rates = np.zeros(n, vb.DQ_RATES_DTYPE)
    mx = (1 << bd) - 1
    items, off = [], 0
    for i, r in enumerate(recs):
        resi = r['resi']
        h, w = resi.shape
        pred = np.clip(rng.integers(0, mx + 1, (h, w)), np.maximum(0, -resi), np.minimum(mx, mx - resi)).astype(np.int16)
        org = np.clip(pred.astype(np.int32) + resi, 0, mx).astype(np.int16)
        cx, cy = 64 * (i % cols), 64 * (i // cols)
        orig[cy:cy + h, cx:cx + w] = org
        j = jobs[i]
        j['x'], j['y'], j['log2w'], j['log2h'], j['mts_idx'] = cx, cy, w.bit_length() - 1, h.bit_length() - 1, 1
        j['flags'] = vb.TU_QUANT | vb.TU_RDOQ_TS
        j['qp_per'], j['qp_rem'], j['offset'], j['rate_idx'], j['lambda'] = r['per'], r['rem'], off, i, r['lambda']
        rates[i] = O.dq_rates_from_flat(None, r['rates'])
        items.append(dict(rec=r, pred=pred, org=org, off=off))
        off += w * h
    return orig, jobs, np.concatenate([r['resi'].ravel() for r in recs]), np.concatenate([it['pred'].ravel() for it in items]), rates, items


def check_rdoq_outputs(items, bd, out):
    errs = []
    for i, it in enumerate(items):
        r = it['rec']
        h, w = r['resi'].shape
        sl = slice(it['off'], it['off'] + w * h)
        res = out['results'][i]
        if not np.array_equal(out['level'][sl].reshape(h, w), r['level']):
            errs.append('job %d %dx%d: levels differ (%d positions)' % (i, w, h, int((out['level'][sl].reshape(h, w) != r['level']).sum())))
        if int(res['abs_sum_level']) != r['abs_sum']:
            errs.append('job %d: abs sum %d != %d' % (i, res['abs_sum_level'], r['abs_sum']))
        deq = O.dequant(r['level'], bd, r['per'], r['rem'], True)
        resi = O.inv_transform(deq, bd, 1)
        reco, sse = O.reconstruct_sse(it['org'], it['pred'], resi, bd)
        if not np.array_equal(out['reco'][sl].reshape(h, w), reco):
            errs.append('job %d %dx%d: reconstruction differs' % (i, w, h))
        if int(res['sse']) != int(sse):
            errs.append('job %d: sse %d != %d' % (i, res['sse'], sse))
    return errs


def random_rdoq_case(rng, bd, n_per_shape):
    import vvc_intra_b200 as vb
    mx = (1 << bd) - 1
    items = []
    for lw in range(2, 6):
        for lh in range(2, 6):
            for _ in range(n_per_shape):
                w, h = 1 << lw, 1 << lh
                a = min(int(rng.choice([1, 4, 20, 120, mx])), mx)
                pred = rng.integers(0, mx + 1, (h, w))
                org = np.clip(pred + rng.integers(-a, a + 1, (h, w)), 0, mx)
                qp = int(rng.integers(4, 52)) + 6 * (bd - 8)
                items.append(dict(pred=pred.astype(np.int16), org=org.astype(np.int16), resi=(org - pred).astype(np.int16), qp=qp,
                                  lam=float(rng.uniform(0.3, 3.0) * 0.57 * 2.0 ** ((qp - 6 * (bd - 8) - 12) / 3.0))))
    n = len(items)
    cols = 16
    orig = np.zeros((64 * ((n + cols - 1) // cols + 1), 64 * cols), np.int16)
    jobs = np.zeros(n, vb.TU_JOB_DTYPE)
    n_rates = 5
    rates = np.zeros(n_rates, vb.DQ_RATES_DTYPE)
    for name in rates.dtype.names:
        rates[name] = rng.integers(300, 140000, rates[name].shape)
    off = 0
    for i, it in enumerate(items):
        h, w = it['resi'].shape
        cx, cy = 64 * (i % cols), 64 * (i // cols)
        orig[cy:cy + h, cx:cx + w] = it['org']
        j = jobs[i]
        j['x'], j['y'], j['log2w'], j['log2h'], j['mts_idx'] = cx, cy, w.bit_length() - 1, h.bit_length() - 1, 1
        j['flags'] = vb.TU_QUANT | vb.TU_RDOQ_TS
        j['qp_per'], j['qp_rem'], j['offset'], j['rate_idx'], j['lambda'] = it['qp'] // 6, it['qp'] % 6, off, i % n_rates, it['lam']
        it['off'], it['rate'] = off, rates[i % n_rates]
        off += w * h
    return orig, jobs, np.concatenate([it['resi'].ravel() for it in items]), np.concatenate([it['pred'].ravel() for it in items]), rates, items


def oracle_rdoq_chain(items, bd):
    import vvc_intra_b200 as vb
    n = sum(it['resi'].size for it in items)
    out = dict(results=np.zeros(len(items), vb.TU_RESULT_DTYPE), coeff=np.zeros(n, np.int32), level=np.zeros(n, np.int32), reco=np.zeros(n, np.int16))
    for i, it in enumerate(items):
        h, w = it['resi'].shape
        sl = slice(it['off'], it['off'] + w * h)
        co = O.fwd_transform(it['resi'], bd, 1)
        lvl, s = O.rdoq_ts(co, bd, it['qp'], it['lam'], it['rate'])
        res = O.inv_transform(O.dequant(lvl, bd, it['qp'] // 6, it['qp'] % 6, True), bd, 1)
        reco, sse = O.reconstruct_sse(it['org'], it['pred'], res, bd)
        out['coeff'][sl], out['level'][sl], out['reco'][sl] = co.ravel(), lvl.ravel(), reco.ravel()
        out['results'][i] = (O.abs_sum_for_preselection(co, 1), s, sse, it.get('bits_fn', lambda lv: 0)(lvl))
    return out


# ---- residual rate estimation: batches from the reference's 'C' records ------------------------------------
def build_rate_batch(tus):
    """One pricing job per 'C' record: levels, TU geometry, mtsIdx and the two mts_coding switches; every record brings its own
    context-state snapshot.  Returns (jobs, levels_flat, states, records)."""
    import vvc_intra_b200 as vb
    recs = [r for r in tus if r['tag'] == 'C']
    jobs = np.zeros(len(recs), vb.TU_JOB_DTYPE)
    states = np.zeros(len(recs), vb.CTX_STATES_DTYPE)
    off = 0
    for i, r in enumerate(recs):
        j = jobs[i]
        j['log2w'], j['log2h'], j['mts_idx'] = r['w'].bit_length() - 1, r['h'].bit_length() - 1, r['mts']
        j['flags'] = vb.TU_QUANT | vb.TU_RATE | (vb.TU_TS_ALLOWED if r['ts_allowed'] else 0) | (vb.TU_MTS_ALLOWED if r['mts_allowed'] else 0)
        j['offset'], j['rate_idx'] = off, i
        st = np.zeros(174, vb.BIN_MODEL_DTYPE)
        st['state'], st['rate'] = r['states'][:, :2], r['states'][:, 2]
        states[i] = np.frombuffer(st.tobytes(), vb.CTX_STATES_DTYPE)[0]
        off += r['w'] * r['h']
    return jobs, np.concatenate([r['level'].ravel() for r in recs]).astype(np.int32), states, recs
