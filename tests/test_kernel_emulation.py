"""CPU tests of the REAL kernel source (vvc_intra_b200/csrc/vvcb_rmd.cuh) executed on host threads through
tests/host_emul/cuda_emul.h, compared with the reference encoder's golden records and with the oracle.
Test-only: the product library has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as O
import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def emul():
    subprocess.check_call(['make', '-s', '-f', 'tests/host_emul/Makefile'], cwd=ROOT)
    return C.CDLL(os.path.join(ROOT, 'tests/host_emul/libemul_rmd.so'))


def run_emul(lib, orig, reco, bd, visits, pred_of_first=False):
    orig = np.ascontiguousarray(orig, np.int16)
    reco = np.ascontiguousarray(reco, np.int16)
    visits = np.ascontiguousarray(visits, O.VISIT_DTYPE)
    res = np.zeros(len(visits), O.RESULT_DTYPE)
    det = np.zeros(len(visits), O.DETAIL_DTYPE)
    pred = None
    pp = None
    if pred_of_first:
        w, h = 1 << int(visits[0]['log2w']), 1 << int(visits[0]['log2h'])
        pred = np.zeros((O.NUM_SLOTS, h, w), np.int16)
        pp = pred.ctypes.data_as(C.c_void_p)
    rc = lib.emul_rmd_eval(orig.ctypes.data_as(C.c_void_p), reco.ctypes.data_as(C.c_void_p), orig.shape[1], bd, 128,
                           visits.ctypes.data_as(C.c_void_p), len(visits), res.ctypes.data_as(C.c_void_p),
                           det.ctypes.data_as(C.c_void_p), pp)
    assert rc == 0
    return res, det, pred


def brief_matches(brief, full):
    """vvcb_rmd_brief against vvcb_rmd_result: counts, final list (its first n_rd entries are the RD list) and Hadamard list as mode codes."""
    code = lambda m: m['mode'].astype(np.uint16) | (m['mrl'].astype(np.uint16) << 8) | (m['mip'].astype(np.uint16) << 15)
    for b, r in zip(brief, full):
        if (int(b['n_rd']), int(b['n_had']), int(b['n_final'])) != (int(r['n_rd']), int(r['n_had']), int(r['n_final'])):
            return False
        nf, nh = int(r['n_final']), int(r['n_had'])
        if not np.array_equal(b['final_mode'][:nf], code(r['final_mode'][:nf])) or b['final_mode'][nf:].any():
            return False
        if not np.array_equal(b['final_mode'][:int(r['n_rd'])], code(r['rd_mode'][:int(r['n_rd'])])):
            return False
        if not np.array_equal(b['had_mode'][:nh], code(r['had_mode'][:nh])) or b['had_mode'][nh:].any():
            return False
    return True


def pick(visits, per_shape):
    seen, out = {}, []
    for v in visits:
        k = (v['head']['w'], v['head']['h'])
        if seen.get(k, 0) < per_shape:
            seen[k] = seen.get(k, 0) + 1
            out.append(v)
    return out


@pytest.mark.parametrize('name', ['ref_8b_128x64_qp32', 'ref_10b_192x128_qp27'])
def test_emulated_kernels_match_reference(emul, name):
    visits, _ = G.load_fixture(name)
    sel = pick(visits, 2)
    orig, reco, arr = G.build_atlas(sel)
    res, det, _ = run_emul(emul, orig, reco, sel[0]['head']['bd'], arr)
    errs = []
    for v, r, d in zip(sel, res, det):
        errs += G.check_visit_against_reference(v, r, d)
    assert not errs, errs[:5]
    # and bit-identical to the oracle on every slot, including the ones the reference never evaluated
    ora, odet = O.rmd_batch(orig, reco, sel[0]['head']['bd'], 128, arr)
    assert np.array_equal(det['sad'], odet['sad']) and np.array_equal(det['satd'], odet['satd'])
    assert res.tobytes() == ora.tobytes() and det.tobytes() == odet.tobytes()


@pytest.mark.parametrize('bd,seed', [(8, 5), (10, 6)])
def test_emulated_kernels_match_oracle_on_random_visits(emul, bd, seed):
    """Ragged availability, NO_MRL / NO_MIP flags, random MPM lists (with and without DC), random rates."""
    rng = np.random.default_rng(seed)
    orig, reco, arr = G.random_case(rng, bd, 4, plane=(256, 512))
    res, det, _ = run_emul(emul, orig, reco, bd, arr)
    ora, odet = O.rmd_batch(orig, reco, bd, 128, arr)
    # without detail tables only min(2 * SAD, SATD) is handed from the evaluation kernels to the list kernel: same lists
    res2 = np.zeros(len(arr), O.RESULT_DTYPE)
    assert emul.emul_rmd_eval(np.ascontiguousarray(orig, np.int16).ctypes.data_as(C.c_void_p), np.ascontiguousarray(reco, np.int16).ctypes.data_as(C.c_void_p), orig.shape[1], bd, 128,
                              arr.ctypes.data_as(C.c_void_p), len(arr), res2.ctypes.data_as(C.c_void_p), None, None) == 0
    assert res2.tobytes() == ora.tobytes()
    # the brief records (vvcb_rmd_eval_brief) carry the same lists as mode codes
    import vvc_intra_b200 as vb
    brief = np.zeros(len(arr), vb.BRIEF_DTYPE)
    assert emul.emul_rmd_brief(arr.ctypes.data_as(C.c_void_p), len(arr), 128, odet.ctypes.data_as(C.c_void_p), brief.ctypes.data_as(C.c_void_p)) == 0
    assert brief_matches(brief, ora)
    bad = [i for i in range(len(arr)) if res[i].tobytes() != ora[i].tobytes() or det[i].tobytes() != odet[i].tobytes()]
    assert not bad, (len(bad), arr[bad[0]])


def test_emulated_packed_items_ragged_tails_and_ctu_rows(emul):
    """Packed work items (eight visits of one small shape per warp item, slot-major task list): 19 visits of each of the eight small
    shapes, i.e. items of 8, 8 and 3 visits, a third of them on a CTU row boundary (no multi-reference-line slots: the visits of one
    item have different slot counts), mixed NO_MRL / NO_MIP flags and ragged availability -- against the oracle."""
    rng = np.random.default_rng(77)
    bd = 10
    orig, reco, arr = G.random_case(rng, bd, 19, plane=(512, 512))
    small = [(2, 2), (3, 2), (2, 3), (3, 3), (4, 2), (2, 4), (4, 3), (3, 4)]
    arr = arr[np.isin(arr['log2w'] * 8 + arr['log2h'], [lw * 8 + lh for lw, lh in small])].copy()
    assert len(arr) == 19 * 8
    on_row = rng.random(len(arr)) < 0.33
    arr['y'][on_row] = 128 * rng.integers(1, 3, int(on_row.sum()))
    res, det, _ = run_emul(emul, orig, reco, bd, arr)
    ora, odet = O.rmd_batch(orig, reco, bd, 128, arr)
    # without detail tables only min(2 * SAD, SATD) is handed from the evaluation kernels to the list kernel: same lists
    res2 = np.zeros(len(arr), O.RESULT_DTYPE)
    assert emul.emul_rmd_eval(np.ascontiguousarray(orig, np.int16).ctypes.data_as(C.c_void_p), np.ascontiguousarray(reco, np.int16).ctypes.data_as(C.c_void_p), orig.shape[1], bd, 128,
                              arr.ctypes.data_as(C.c_void_p), len(arr), res2.ctypes.data_as(C.c_void_p), None, None) == 0
    assert res2.tobytes() == ora.tobytes()
    # the brief records (vvcb_rmd_eval_brief) carry the same lists as mode codes
    import vvc_intra_b200 as vb
    brief = np.zeros(len(arr), vb.BRIEF_DTYPE)
    assert emul.emul_rmd_brief(arr.ctypes.data_as(C.c_void_p), len(arr), 128, odet.ctypes.data_as(C.c_void_p), brief.ctypes.data_as(C.c_void_p)) == 0
    assert brief_matches(brief, ora)
    bad = [i for i in range(len(arr)) if res[i].tobytes() != ora[i].tobytes() or det[i].tobytes() != odet[i].tobytes()]
    assert not bad, (len(bad), arr[bad[0]])
    assert (det['sad'][on_row][:, 67:77] == 0xFFFFFFFF).all()          # no MRL evaluations on a CTU row boundary


# ---- TU coding kernel (vvcb_tu.cuh) ------------------------------------------------------------------------
def run_emul_tu(lib, orig, bd, jobs, resi, pred, rates=None, states=None):
    import vvc_intra_b200 as vb
    orig = np.ascontiguousarray(orig, np.int16)
    jobs = np.ascontiguousarray(jobs, vb.TU_JOB_DTYPE)
    resi = np.ascontiguousarray(resi, np.int16)
    pred = np.ascontiguousarray(pred, np.int16)
    lib.emul_tu_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    nr = len(rates) if rates is not None else (len(states) if states is not None else 0)
    rates = None if rates is None else np.ascontiguousarray(rates, vb.DQ_RATES_DTYPE)
    states = None if states is None else np.ascontiguousarray(states, vb.CTX_STATES_DTYPE)
    out = dict(results=np.zeros(len(jobs), vb.TU_RESULT_DTYPE), coeff=np.zeros(resi.size, np.int32), level=np.zeros(resi.size, np.int32),
               reco=np.zeros(resi.size, np.int16))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emul_tu_eval(p(orig), orig.shape[1], bd, p(jobs), len(jobs), p(resi), p(pred), resi.size, p(rates) if rates is not None else None,
                          p(states) if states is not None else None, nr,
                          p(out['coeff']), p(out['level']), p(out['reco']), p(out['results']))
    assert rc == 0
    return out


@pytest.mark.parametrize('name,bd', [('ref_8b_128x64_qp32', 8), ('ref_10b_192x128_qp27', 10), ('ref_10b_64x64_qp32_scalarq', 10)])
def test_emulated_tu_kernel_matches_reference(emul, name, bd):
    _, tus = G.load_fixture(name)
    orig, jobs, resi, pred, items = G.build_tu_batch(tus, bd, max_jobs=160)
    assert len(items) > 20
    out = run_emul_tu(emul, orig, bd, jobs, resi, pred)
    errs = G.check_tu_outputs(items, bd, out)
    assert not errs, errs[:5]


@pytest.mark.parametrize('bd,seed', [(8, 21), (10, 22)])
def test_emulated_tu_kernel_matches_oracle_on_random_blocks(emul, bd, seed):
    rng = np.random.default_rng(seed)
    orig, jobs, resi, pred, items = G.random_tu_case(rng, bd, 1)
    out = run_emul_tu(emul, orig, bd, jobs, resi, pred)
    exp = G.oracle_tu_chain(items, bd)
    for k in ('coeff', 'level', 'reco'):
        assert np.array_equal(out[k], exp[k]), k
    assert out['results'].tobytes() == exp['results'].tobytes()


# ---- texture-measure kernels (vvcb_feat.cuh) -----------------------------------------------------------------
def test_emulated_ctu_hads_kernel_matches_reference_records(emul):
    _, recs = G.load_fixture('ref_10b_200x136_ctuhad')
    hs = [r for r in recs if r['tag'] == 'H']
    pic = np.zeros((136, 256), np.int16)            # pitch 256 > width 200
    pic[:128, :128], pic[:128, 128:200], pic[128:, :128], pic[128:, 128:200] = [r['org'] for r in hs]
    out = np.zeros(4, np.int32)
    emul.emul_ctu_hads(pic.ctypes.data_as(C.c_void_p), 256, 200, 136, 128, out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == [r['result'] for r in hs]
    rng = np.random.default_rng(3)
    pic = rng.integers(0, 1024, (72, 192)).astype(np.int16)
    out = np.zeros(6, np.int32)
    emul.emul_ctu_hads(pic.ctypes.data_as(C.c_void_p), 192, 192, 72, 64, out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == O.ctu_hads_islice(pic, ctu=64).tolist()


@pytest.mark.parametrize('seed,mx', [(41, 256), (42, 1024), (43, 40)])
def test_emulated_features_kernel_matches_oracle(emul, seed, mx):
    from test_oracle_features import random_feature_jobs
    rng = np.random.default_rng(seed)
    H, W = 128, 256
    pic = rng.integers(0, mx, (H, W)).astype(np.int16)
    pic[:, 128:] = (pic[:, 128:] // 8) * 8 % 256          # smoother half: ties in the rounded gradient mean
    jobs = random_feature_jobs(rng, H, W, 120)
    out = np.zeros(len(jobs), O.FEAT_RESULT_DTYPE)
    emul.emul_features_eval(pic.ctypes.data_as(C.c_void_p), W, jobs.ctypes.data_as(C.c_void_p), len(jobs), out.ctypes.data_as(C.c_void_p))
    exp = O.features_batch(pic, jobs)
    bad = [i for i in range(len(jobs)) if out[i].tobytes() != exp[i].tobytes()]
    assert not bad, (jobs[bad[0]]['cu'], out[bad[0]]['f'].tolist(), exp[bad[0]]['f'].tolist())


# ---- dependent quantisation kernel (vvcb_dq.cuh) --------------------------------------------------------------
@pytest.mark.parametrize('name,bd', [('ref_10b_128x128_qp27_depquant', 10), ('ref_8b_128x64_qp37_depquant', 8)])
def test_emulated_dq_kernel_matches_reference(emul, name, bd):
    _, tus = G.load_fixture(name)
    orig, jobs, resi, pred, rates, items = G.build_dq_batch(tus, bd)
    assert len(items) > 80
    out = run_emul_tu(emul, orig, bd, jobs, resi, pred, rates)
    errs = G.check_dq_outputs(items, bd, out)
    assert not errs, (len(errs), errs[:6])


@pytest.mark.parametrize('bd,seed', [(8, 61), (10, 62)])
def test_emulated_dq_kernel_matches_oracle_on_random_blocks(emul, bd, seed):
    rng = np.random.default_rng(seed)
    orig, jobs, resi, pred, rates, items = G.random_dq_case(rng, bd, 1)
    out = run_emul_tu(emul, orig, bd, jobs, resi, pred, rates)
    exp = G.oracle_dq_chain(items, bd)
    for k in ('coeff', 'level', 'reco'):
        bad = [i for i, it in enumerate(items) if not np.array_equal(out[k][it['off']:it['off'] + it['resi'].size], exp[k][it['off']:it['off'] + it['resi'].size])]
        assert not bad, (k, len(bad), [(items[i]['resi'].shape, items[i]['mts'], items[i]['qp'], items[i]['lfnst']) for i in bad[:5]])
    assert out['results'].tobytes() == exp['results'].tobytes()


# ---- prediction + residual of TU jobs (tu_pred_kernel) -----------------------------------------------------------
@pytest.mark.parametrize('bd,seed', [(8, 81), (10, 82)])
def test_emulated_tu_prediction_matches_oracle(emul, bd, seed):
    rng = np.random.default_rng(seed)
    orig, reco, visits, src, jobs, n_samples, rates, items = G.pred_tu_case(rng, bd, 2)
    pred = np.zeros(n_samples, np.int16)
    resi = np.zeros(n_samples, np.int16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    orig = np.ascontiguousarray(orig, np.int16)
    reco = np.ascontiguousarray(reco, np.int16)
    emul.emul_tu_pred(p(orig), p(reco), orig.shape[1], bd, 128, p(visits), p(src), p(jobs), len(jobs), p(pred), p(resi))
    kinds = set()
    for s, it in zip(src, items):
        sl = slice(it['off'], it['off'] + it['pred'].size)
        assert np.array_equal(pred[sl].reshape(it['pred'].shape), it['pred']), (it['pred'].shape, int(s['slot']))
        assert np.array_equal(resi[sl].reshape(it['pred'].shape), it['resi']), (it['pred'].shape, int(s['slot']))
        kinds.add('mip' if s['slot'] >= O.SLOT_MIP else 'mrl' if s['slot'] >= O.SLOT_MRL1 else 'reg')
    assert kinds == {'mip', 'mrl', 'reg'}


# ---- RDOQ of transform-skip TUs (rdoq_ts_kernel) -----------------------------------------------------------------
@pytest.mark.parametrize('name,bd', [('ref_10b_128x128_qp27_rdoqts', 10), ('ref_8b_128x64_qp37_rdoqts', 8)])
def test_emulated_rdoq_ts_kernel_matches_reference(emul, name, bd):
    _, tus = G.load_fixture(name)
    orig, jobs, resi, pred, rates, items = G.build_rdoq_batch(tus, bd)
    assert len(items) > 100
    out = run_emul_tu(emul, orig, bd, jobs, resi, pred, rates)
    errs = G.check_rdoq_outputs(items, bd, out)
    assert not errs, (len(errs), errs[:6])


@pytest.mark.parametrize('bd,seed', [(8, 101), (10, 102)])
def test_emulated_rdoq_ts_kernel_matches_oracle_on_random_blocks(emul, bd, seed):
    rng = np.random.default_rng(seed)
    orig, jobs, resi, pred, rates, items = G.random_rdoq_case(rng, bd, 3)
    out = run_emul_tu(emul, orig, bd, jobs, resi, pred, rates)
    exp = G.oracle_rdoq_chain(items, bd)
    for k in ('coeff', 'level', 'reco'):
        bad = [i for i, it in enumerate(items) if not np.array_equal(out[k][it['off']:it['off'] + it['resi'].size], exp[k][it['off']:it['off'] + it['resi'].size])]
        assert not bad, (k, len(bad), [(items[i]['resi'].shape, items[i]['qp']) for i in bad[:5]])
    assert out['results'].tobytes() == exp['results'].tobytes()
    assert (exp['results']['abs_sum_level'] > 0).sum() > len(items) // 3


# ---- LFNST inside the TU kernel -------------------------------------------------------------------------------------
@pytest.mark.parametrize('name,bd', [('ref_10b_128x128_qp27_lfnst', 10), ('ref_8b_128x64_qp32_lfnst', 8)])
def test_emulated_lfnst_matches_reference(emul, name, bd):
    """Primary transform restricted to the LFNST region, forward LFNST, dependent quantisation from scan position 7 / 15,
    dequantisation, inverse LFNST, inverse primary, reconstruction: against the reference's 'F' records and the oracle chain
    (whose inverse half is pinned by the 'J' records)."""
    _, tus = G.load_fixture(name)
    orig, jobs, resi, pred, rates, items = G.build_dq_batch(tus, bd, tag='F')
    assert len(items) > 50
    out = run_emul_tu(emul, orig, bd, jobs, resi, pred, rates)
    errs = G.check_dq_outputs(items, bd, out)
    assert not errs, (len(errs), errs[:6])


# ---- residual rate estimation (rate_kernel) ----------------------------------------------------------------------------
def run_emul_rate(lib, jobs, levels, states):
    import vvc_intra_b200 as vb
    jobs = np.ascontiguousarray(jobs, vb.TU_JOB_DTYPE)
    levels = np.ascontiguousarray(levels, np.int32)
    states = np.ascontiguousarray(states, vb.CTX_STATES_DTYPE)
    res = np.zeros(len(jobs), vb.TU_RESULT_DTYPE)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.emul_residual_bits(p(jobs), len(jobs), p(levels), p(states), 1, p(res))
    return res['frac_bits']


@pytest.mark.parametrize('name', ['ref_10b_128x128_qp27_resbits', 'ref_8b_128x64_qp22_resbits'])
def test_emulated_rate_kernel_matches_reference(emul, name):
    _, tus = G.load_fixture(name)
    jobs, levels, states, recs = G.build_rate_batch(tus)
    assert len(recs) > 150
    got = run_emul_rate(emul, jobs, levels, states)
    bad = [(r['w'], r['h'], r['mts'], int(g), r['bits']) for g, r in zip(got, recs) if int(g) != r['bits']]
    assert not bad, (len(bad), bad[:5])


def test_mode_params_match_the_reference_function(emul):
    """make_mode_param (vvcb_core.cuh, the source the library's parameter ROM is built from) against the UNMODIFIED reference's compiled
    IntraPrediction::initPredIntraParams (CL/IntraPrediction.cpp:487-618) for 17 shapes x 67 modes x reference lines 0 / 1 / 3:
    tests/golden/intra_params.txt.gz, written by oracle/dump_intra_params.cpp through oracle/_ref/libvtmref.a."""
    import gzip
    emul.emul_mode_param.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    out = np.zeros(7, np.int32)
    n = 0
    shapes = set()
    for line in gzip.open(os.path.join(ROOT, 'tests/golden/intra_params.txt.gz'), 'rt'):
        if line.startswith('#') or line.startswith('ISP'):    # the ISP part of the table: tests/test_isp_plan.py
            continue
        head, _, tail = line.partition('|')
        w, h, mode, mrl = (int(v) for v in head.split())
        is_ver, ref_filter, interp, pdpc, angle, inv_angle, scale = (int(v) for v in tail.split())
        emul.emul_mode_param(w, h, mode, mrl, out.ctypes.data_as(C.c_void_p))
        key = (w, h, mode, mrl)
        assert (out[0], out[1], out[2], out[3]) == (is_ver, ref_filter, interp, pdpc), key
        if mode > 1:                              # the reference leaves the angle fields untouched for planar / DC
            assert (out[4], out[5]) == (angle, inv_angle), key
            if angle > 0 and pdpc:                # angularScale is only written for positive angles and only read with PDPC
                assert out[6] == scale, key
        n += 1
        shapes.add((w, h))
    assert len(shapes) == 17 and n == 17 * (67 + 66 + 66)
