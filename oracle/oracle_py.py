"""ctypes binding of the plain-C oracle (oracle/libvvc_oracle.so).

TEST INFRASTRUCTURE: import only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

NUM_SLOTS, MAX_LIST = 112, 16
SLOT_MRL1, SLOT_MRL3, SLOT_MIP = 67, 72, 77
SAT_NONE = 0xFFFFFFFF


class Rates(C.Structure):
    _fields_ = [('mip_flag', C.c_uint32 * 2), ('mrl_bin0', C.c_uint32 * 2), ('mrl_bin1', C.c_uint32 * 2),
                ('isp_bin0_0', C.c_uint32), ('mpm_flag', C.c_uint32 * 2), ('planar_flag', C.c_uint32 * 2)]


class RmdVisit(C.Structure):
    _fields_ = [('x', C.c_int16), ('y', C.c_int16), ('log2w', C.c_uint8), ('log2h', C.c_uint8),
                ('avail_al', C.c_uint8), ('n_above', C.c_uint8), ('n_above_right', C.c_uint8),
                ('n_left', C.c_uint8), ('n_below_left', C.c_uint8), ('flags', C.c_uint8),
                ('mpm', C.c_uint8 * 6), ('num_mpm_cand', C.c_uint8), ('pad', C.c_uint8 * 3),
                ('rates', Rates), ('sqrt_lambda', C.c_double)]


class Mode(C.Structure):
    _fields_ = [('mip', C.c_uint8), ('mrl', C.c_uint8), ('mode', C.c_uint8), ('pad', C.c_uint8)]


MAX_HAD_LIST = 8


class RmdResult(C.Structure):
    _fields_ = [('n_rd', C.c_int32), ('n_had', C.c_int32), ('n_final', C.c_int32), ('pad', C.c_int32),
                ('rd_mode', Mode * MAX_LIST), ('rd_cost', C.c_double * MAX_LIST),
                ('had_mode', Mode * MAX_HAD_LIST), ('had_cost', C.c_double * MAX_HAD_LIST),
                ('final_mode', Mode * MAX_LIST)]


class RmdDetail(C.Structure):
    _fields_ = [('sad', C.c_uint32 * NUM_SLOTS), ('satd', C.c_uint32 * NUM_SLOTS), ('n_reg', C.c_int32), ('n_reg_had', C.c_int32),
                ('reg_mode', Mode * MAX_LIST), ('reg_cost', C.c_double * MAX_LIST),
                ('reg_had_mode', Mode * MAX_HAD_LIST), ('reg_had_cost', C.c_double * MAX_HAD_LIST)]


class Ipa(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('is_ver', 'mrl', 'ref_filter', 'interp', 'pdpc', 'angle', 'inv_angle', 'ang_scale')]


VISIT_DTYPE = np.dtype([('x', '<i2'), ('y', '<i2'), ('log2w', 'u1'), ('log2h', 'u1'), ('avail_al', 'u1'),
                        ('n_above', 'u1'), ('n_above_right', 'u1'), ('n_left', 'u1'), ('n_below_left', 'u1'),
                        ('flags', 'u1'), ('mpm', 'u1', 6), ('num_mpm_cand', 'u1'), ('pad', 'u1', 3),
                        ('rates', '<u4', 11), ('sqrt_lambda', '<f8')], align=True)
MODE_DTYPE = np.dtype([('mip', 'u1'), ('mrl', 'u1'), ('mode', 'u1'), ('pad', 'u1')])


RESULT_DTYPE = np.dtype([('n_rd', '<i4'), ('n_had', '<i4'), ('n_final', '<i4'), ('pad', '<i4'),
                         ('rd_mode', MODE_DTYPE, MAX_LIST), ('rd_cost', '<f8', MAX_LIST),
                         ('had_mode', MODE_DTYPE, MAX_HAD_LIST), ('had_cost', '<f8', MAX_HAD_LIST),
                         ('final_mode', MODE_DTYPE, MAX_LIST)], align=True)
DETAIL_DTYPE = np.dtype([('sad', '<u4', NUM_SLOTS), ('satd', '<u4', NUM_SLOTS), ('n_reg', '<i4'), ('n_reg_had', '<i4'),
                         ('reg_mode', MODE_DTYPE, MAX_LIST), ('reg_cost', '<f8', MAX_LIST),
                         ('reg_had_mode', MODE_DTYPE, MAX_HAD_LIST), ('reg_had_cost', '<f8', MAX_HAD_LIST)], align=True)
assert VISIT_DTYPE.itemsize == C.sizeof(RmdVisit), (VISIT_DTYPE.itemsize, C.sizeof(RmdVisit))
assert RESULT_DTYPE.itemsize == C.sizeof(RmdResult), (RESULT_DTYPE.itemsize, C.sizeof(RmdResult))
assert DETAIL_DTYPE.itemsize == C.sizeof(RmdDetail), (DETAIL_DTYPE.itemsize, C.sizeof(RmdDetail))

_lib = None
_p16 = C.POINTER(C.c_int16)


def build():
    subprocess.check_call(['make', '-s', '-f', 'oracle/Makefile'], cwd=_ROOT)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, 'libvvc_oracle.so')
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_sad.restype = C.c_uint64
        L.orc_satd.restype = C.c_uint64
        L.orc_mode_bits.restype = C.c_uint64
        L.orc_fnv1a.restype = C.c_uint64
        _lib = L
    return _lib


def _a16(a):
    a = np.ascontiguousarray(a, dtype=np.int16)
    return a, a.ctypes.data_as(_p16)


def ref_fill(reco, x, y, w, h, mrl, bd, avail_al, n_above, n_above_right, n_left, n_below_left):
    reco, pr = _a16(reco)
    top = np.zeros(2 * w + 1 + mrl, np.int16)
    left = np.zeros(2 * h + 1 + mrl, np.int16)
    lib().orc_ref_fill(pr, reco.shape[1], x, y, w, h, mrl, bd, avail_al, n_above, n_above_right, n_left, n_below_left,
                       top.ctypes.data_as(_p16), left.ctypes.data_as(_p16))
    return top, left


def ref_filter(top, left, w, h, mrl):
    top, pt = _a16(top)
    left, pl = _a16(left)
    ft, fl = np.zeros_like(top), np.zeros_like(left)
    lib().orc_ref_filter(pt, pl, w, h, mrl, ft.ctypes.data_as(_p16), fl.ctypes.data_as(_p16))
    return ft, fl


def ipa_init(w, h, mode, mrl, is_mip=0):
    p = Ipa()
    lib().orc_ipa_init(w, h, mode, mrl, is_mip, C.byref(p))
    return p


def pred_regular(top, left, w, h, bd, mode, ipa):
    top, pt = _a16(top)
    left, pl = _a16(left)
    pred = np.zeros((h, w), np.int16)
    lib().orc_pred_regular(pt, pl, w, h, bd, mode, C.byref(ipa), pred.ctypes.data_as(_p16))
    return pred


def pred_mip(top, left, w, h, bd, mode):
    top, pt = _a16(top)
    left, pl = _a16(left)
    pred = np.zeros((h, w), np.int16)
    lib().orc_pred_mip(pt, pl, w, h, bd, mode, pred.ctypes.data_as(_p16))
    return pred


def sad(org, cur):
    org, po = _a16(org)
    cur, pc = _a16(cur)
    h, w = org.shape
    return lib().orc_sad(po, w, pc, w, w, h)


def satd(org, cur):
    org, po = _a16(org)
    cur, pc = _a16(cur)
    h, w = org.shape
    return lib().orc_satd(po, w, pc, w, w, h)


def fnv1a(a):
    a, p = _a16(a)
    return lib().orc_fnv1a(p, a.size)


def mode_bits(rates11, mpm, w, h, mrl_allowed, mip_enabled, is_mip, mrl, mode):
    r = Rates.from_buffer_copy(np.asarray(rates11, '<u4').tobytes())
    m = (C.c_uint8 * 6)(*[int(v) for v in mpm])
    return lib().orc_mode_bits(C.byref(r), m, w, h, int(mrl_allowed), int(mip_enabled), int(is_mip), mrl, mode)


def intra_mpms(left_dir, above_dir):
    m = (C.c_uint8 * 6)()
    n = C.c_int()
    lib().orc_intra_mpms(left_dir, above_dir, m, C.byref(n))
    return list(m), n.value


def rmd_batch(orig, reco, bd, ctu_size, visits, want_pred=False):
    """visits: numpy array of VISIT_DTYPE.  Returns (results, details) numpy arrays (and preds list)."""
    orig, po = _a16(orig)
    reco, pr = _a16(reco)
    visits = np.ascontiguousarray(visits, dtype=VISIT_DTYPE)
    out = np.zeros(len(visits), RESULT_DTYPE)
    det = np.zeros(len(visits), DETAIL_DTYPE)
    if not want_pred:
        lib().orc_rmd_batch(po, orig.shape[1], pr, reco.shape[1], bd, ctu_size,
                            visits.ctypes.data_as(C.c_void_p), len(visits), out.ctypes.data_as(C.c_void_p),
                            det.ctypes.data_as(C.c_void_p))
        return out, det
    preds = []
    for i in range(len(visits)):
        w, h = 1 << int(visits[i]['log2w']), 1 << int(visits[i]['log2h'])
        p = np.zeros((NUM_SLOTS, h, w), np.int16)
        lib().orc_rmd_visit(po, orig.shape[1], pr, reco.shape[1], bd, ctu_size,
                            C.c_void_p(visits[i:i + 1].ctypes.data), C.c_void_p(out[i:i + 1].ctypes.data),
                            C.c_void_p(det[i:i + 1].ctypes.data), p.ctypes.data_as(_p16))
        preds.append(p)
    return out, det, preds


# ---- TU coding (oracle/vvc_oracle_tr.c) -------------------------------------------------------------------
_p32 = C.POINTER(C.c_int32)


def fwd_transform(resi, bd, mts_idx, lfnst_idx=0):
    resi, pr = _a16(resi)
    h, w = resi.shape
    coeff = np.zeros((h, w), np.int32)
    if mts_idx == 1:
        lib().orc_transform_skip(pr, w, w, h, bd, coeff.ctypes.data_as(_p32))
    else:
        lib().orc_fwd_transform_ex(pr, w, w, h, bd, mts_idx, lfnst_idx, coeff.ctypes.data_as(_p32))
    return coeff


def abs_sum_for_preselection(coeff, mts_idx):
    coeff = np.ascontiguousarray(coeff, np.int32)
    h, w = coeff.shape
    return lib().orc_abs_sum_for_preselection(coeff.ctypes.data_as(_p32), w, h, mts_idx)


def mts_preselect(sums, w, h, max_cand):
    a = (C.c_int * len(sums))(*[int(x) for x in sums])
    sel = (C.c_uint8 * len(sums))()
    lib().orc_mts_preselect(a, len(sums), w, h, max_cand, sel)
    return list(sel)


def quant_scalar(coeff, bd, per, rem, is_ts):
    coeff = np.ascontiguousarray(coeff, np.int32)
    h, w = coeff.shape
    level = np.zeros((h, w), np.int32)
    s = lib().orc_quant_scalar(coeff.ctypes.data_as(_p32), w, h, bd, per, rem, int(is_ts), level.ctypes.data_as(_p32))
    return level, s


def dequant(level, bd, per, rem, is_ts):
    level = np.ascontiguousarray(level, np.int32)
    h, w = level.shape
    coeff = np.zeros((h, w), np.int32)
    lib().orc_dequant(level.ctypes.data_as(_p32), w, h, bd, per, rem, int(is_ts), coeff.ctypes.data_as(_p32))
    return coeff


def inv_transform(coeff, bd, mts_idx):
    coeff = np.ascontiguousarray(coeff, np.int32)
    h, w = coeff.shape
    resi = np.zeros((h, w), np.int16)
    if mts_idx == 1:
        lib().orc_inv_transform_skip(coeff.ctypes.data_as(_p32), w, h, bd, resi.ctypes.data_as(_p16), w)
    else:
        lib().orc_inv_transform(coeff.ctypes.data_as(_p32), w, h, bd, mts_idx, resi.ctypes.data_as(_p16), w)
    return resi


def reconstruct_sse(org, pred, resi, bd):
    org, po = _a16(org)
    pred, pp = _a16(pred)
    resi, pr = _a16(resi)
    h, w = org.shape
    reco = np.zeros((h, w), np.int16)
    lib().orc_reconstruct_sse.restype = C.c_uint64
    sse = lib().orc_reconstruct_sse(po, w, pp, pr, w, h, bd, reco.ctypes.data_as(_p16))
    return reco, sse


# ---- texture measures (oracle/vvc_oracle_feat.c) ------------------------------------------------------------
FEAT_CU_DTYPE = np.dtype([('x', '<i2'), ('y', '<i2'), ('w', 'u1'), ('h', 'u1'), ('qt_depth', 'u1'), ('mt_depth', 'u1')])
FEAT_JOB_DTYPE = np.dtype([('cu', FEAT_CU_DTYPE), ('n_neighbours', 'u1'), ('pad', 'u1', 7), ('nb', FEAT_CU_DTYPE, 5)])
FEAT_RESULT_DTYPE = np.dtype([('f', '<i4', 27), ('valid', '<i4')])
assert FEAT_JOB_DTYPE.itemsize == 56 and FEAT_RESULT_DTYPE.itemsize == 112


def ctu_hads_islice(orig, ctu=128):
    orig, po = _a16(orig)
    h, w = orig.shape
    out = np.zeros(((h + ctu - 1) // ctu) * ((w + ctu - 1) // ctu), np.int32)
    lib().orc_ctu_hads_islice(po, w, w, h, ctu, out.ctypes.data_as(_p32))
    return out


def features_batch(orig, jobs):
    orig, po = _a16(orig)
    jobs = np.ascontiguousarray(jobs, FEAT_JOB_DTYPE)
    out = np.zeros(len(jobs), FEAT_RESULT_DTYPE)
    lib().orc_features_batch(po, orig.shape[1], C.c_void_p(jobs.ctypes.data), len(jobs), C.c_void_p(out.ctypes.data))
    return out


# ---- dependent quantisation (oracle/vvc_oracle_dq.c) -------------------------------------------------------
DQ_RATES_DTYPE = np.dtype([('sig_sbb', '<u4', (2, 2)), ('sig', '<u4', (3, 12, 2)), ('par', '<u4', (21, 2)), ('gt1', '<u4', (21, 2)),
                           ('gt2', '<u4', (21, 2)), ('last_x', '<u4', (20, 2)), ('last_y', '<u4', (20, 2)),
                           ('ts_sig_sbb', '<u4', (3, 2)), ('ts_sig', '<u4', (3, 2)), ('ts_par', '<u4', (1, 2)), ('ts_gtx', '<u4', (5, 2)),
                           ('ts_lrg1', '<u4', (4, 2)), ('ts_sign', '<u4', (6, 2))])
assert DQ_RATES_DTYPE.itemsize == 4 * 2 * (2 + 36 + 63 + 40 + 22)


def dq_rates_from_flat(flat, ts_flat=None):
    """The 'D' record's flat context-price vector (282 words) and / or the 'T' record's (44 words) -> one vvcb_dq_rates struct."""
    buf = np.zeros(DQ_RATES_DTYPE.itemsize // 4, '<u4')
    if flat is not None:
        buf[:282] = flat
    if ts_flat is not None:
        buf[282:] = ts_flat
    return np.frombuffer(buf.tobytes(), DQ_RATES_DTYPE)[0]


def dep_quant(coeff, bd, mts_idx, lfnst_idx, qp, lam, rates, cbf_delta_bits):
    coeff = np.ascontiguousarray(coeff, np.int32)
    h, w = coeff.shape
    level = np.zeros((h, w), np.int32)
    r = np.ascontiguousarray(np.array([rates], DQ_RATES_DTYPE))
    s = lib().orc_dep_quant(coeff.ctypes.data_as(_p32), w, h, bd, mts_idx, lfnst_idx, qp, C.c_double(lam), C.c_void_p(r.ctypes.data),
                            cbf_delta_bits, level.ctypes.data_as(_p32))
    return level, s


def dep_dequant(level, bd, qp):
    level = np.ascontiguousarray(level, np.int32)
    h, w = level.shape
    coeff = np.zeros((h, w), np.int32)
    lib().orc_dep_dequant(level.ctypes.data_as(_p32), w, h, bd, qp, coeff.ctypes.data_as(_p32))
    return coeff


def rdoq_ts(coeff, bd, qp, lam, rates):
    """QuantRDOQ::xRateDistOptQuantTS.  coeff: transform-skip coefficients (residual << transformShift)."""
    coeff = np.ascontiguousarray(coeff, np.int32)
    h, w = coeff.shape
    level = np.zeros((h, w), np.int32)
    r = np.ascontiguousarray(np.array([rates], DQ_RATES_DTYPE))
    s = lib().orc_rdoq_ts(coeff.ctypes.data_as(_p32), w, h, bd, qp, C.c_double(lam), C.c_void_p(r.ctypes.data), level.ctypes.data_as(_p32))
    return level, s


def fwd_lfnst(coeff, intra_mode, lfnst_idx):
    c = np.ascontiguousarray(coeff, np.int32).copy()
    h, w = c.shape
    lib().orc_fwd_lfnst(c.ctypes.data_as(_p32), w, h, intra_mode, lfnst_idx)
    return c


def inv_lfnst(coeff, intra_mode, lfnst_idx):
    c = np.ascontiguousarray(coeff, np.int32).copy()
    h, w = c.shape
    lib().orc_inv_lfnst(c.ctypes.data_as(_p32), w, h, intra_mode, lfnst_idx)
    return c


# ---- residual rate estimation (oracle/vvc_oracle_rate.c) ---------------------------------------------------------
BIN_MODEL_DTYPE = np.dtype([('state', '<u2', 2), ('rate', 'u1'), ('pad', 'u1')])
CTX_STATES_DTYPE = np.dtype([('mts_idx', BIN_MODEL_DTYPE, 11), ('sig_sbb', BIN_MODEL_DTYPE, 2), ('sig', BIN_MODEL_DTYPE, (3, 12)), ('par', BIN_MODEL_DTYPE, 21),
                             ('gt1', BIN_MODEL_DTYPE, 21), ('gt2', BIN_MODEL_DTYPE, 21), ('last_x', BIN_MODEL_DTYPE, 20), ('last_y', BIN_MODEL_DTYPE, 20),
                             ('ts_sig_sbb', BIN_MODEL_DTYPE, 3), ('ts_sig', BIN_MODEL_DTYPE, 3), ('ts_par', BIN_MODEL_DTYPE, 1), ('ts_gtx', BIN_MODEL_DTYPE, 5),
                             ('ts_lrg1', BIN_MODEL_DTYPE, 4), ('ts_sign', BIN_MODEL_DTYPE, 6)])
assert CTX_STATES_DTYPE.itemsize == 174 * 6


def ctx_states_from_record(states):
    """The 'C' record's (174, 3) array of (state0, state1, rate) -> one vvcb_ctx_states struct."""
    flat = np.zeros(174, BIN_MODEL_DTYPE)
    flat['state'] = states[:, :2]
    flat['rate'] = states[:, 2]
    return np.frombuffer(flat.tobytes(), CTX_STATES_DTYPE)[0]


def residual_bits(level, mts_idx, ts_allowed, mts_allowed, dep_quant, states):
    level = np.ascontiguousarray(level, np.int32)
    h, w = level.shape
    st = np.ascontiguousarray(np.array([states], CTX_STATES_DTYPE))
    lib().orc_residual_bits.restype = C.c_uint64
    return int(lib().orc_residual_bits(level.ctypes.data_as(_p32), w, h, mts_idx, int(ts_allowed), int(mts_allowed), int(dep_quant), C.c_void_p(st.ctypes.data)))
