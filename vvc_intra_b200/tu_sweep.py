"""Host logic of the TU-coding sweep used by bench.py and the GPU tests at full size: one transform-unit job per candidate
CU of the exhaustive sweep (vvc_intra_b200.partition) and per candidate transform, with the residual the rough-mode-decision
winner would leave.  Mirrors what IntraSearch::xIntraCodingTUBlock hands to TrQuant (EL/IntraSearch.cpp:2694-3168): the
residual block, the prediction block, QP, lambda and the estimator's context prices."""
import numpy as np

from .engine import TU_JOB_DTYPE, TU_QUANT, TU_DEPQUANT, DQ_RATES_DTYPE


def lambda_for_qp(qp, bit_depth):
    """Quant::m_dLambda of an intra picture at the shipped configuration (EL/EncSlice.cpp:397-470, all-intra, no hierarchy):
    0.57 * 2^((qp - 12) / 3); the encoder works in the internal-bit-depth QP domain, which cancels in this expression."""
    return 0.57 * 2.0 ** ((qp - 12) / 3.0)


def default_dq_rates(seed=0):
    """Context prices of a freshly initialised estimator are not reproducible without the CABAC tables; the sweep uses a
    fixed pseudo-random snapshot in the range real snapshots show (tests/golden 'D' records: ~2^9 .. 2^18)."""
    rng = np.random.default_rng(seed)
    r = np.zeros(1, DQ_RATES_DTYPE)
    for name in r.dtype.names:
        r[name] = (2.0 ** rng.uniform(9.5, 17.0, r[name].shape)).astype(np.uint32)
    return r


def build_tu_sweep(orig, visits, qp, bit_depth, transforms=(0,), dep_quant=True, max_side=32, smooth=8):
    """orig: (H, W) int16 picture.  visits: sweep visits (x, y, log2w, log2h).  The prediction of a job is the picture
    low-passed over `smooth` x `smooth` cells (what a good intra prediction leaves is the fine texture), so that
    residual = orig - pred has realistic statistics.  Returns (jobs, resi_flat, pred_flat, rates)."""
    H, W = orig.shape
    hh, ww = (H + smooth - 1) // smooth * smooth, (W + smooth - 1) // smooth * smooth
    pad = np.pad(orig.astype(np.int32), ((0, hh - H), (0, ww - W)), mode='edge')
    low = pad.reshape(hh // smooth, smooth, ww // smooth, smooth).mean(axis=(1, 3)).round().astype(np.int32)
    pred_pic = np.clip(low.repeat(smooth, 0).repeat(smooth, 1)[:H, :W], 0, (1 << bit_depth) - 1).astype(np.int16)
    resi_pic = (orig.astype(np.int32) - pred_pic).astype(np.int16)
    sel = visits[(visits['log2w'] <= int(np.log2(max_side))) & (visits['log2h'] <= int(np.log2(max_side)))]
    n = len(sel) * len(transforms)
    jobs = np.zeros(n, TU_JOB_DTYPE)
    sizes = (1 << sel['log2w'].astype(np.int64)) * (1 << sel['log2h'].astype(np.int64))
    sizes = np.repeat(sizes, len(transforms))
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    total = int(sizes.sum())
    for k, name in enumerate(('x', 'y', 'log2w', 'log2h')):
        jobs[name] = np.repeat(sel[name], len(transforms))
    jobs['mts_idx'] = np.tile(np.array(transforms, np.uint8), len(sel))
    jobs['flags'] = TU_QUANT | (TU_DEPQUANT if dep_quant else 0)
    jobs['flags'][jobs['mts_idx'] == 1] = TU_QUANT                      # transform skip never takes the trellis (CL/DepQuant.cpp:1757)
    big = (jobs['mts_idx'] > 0) & ((jobs['log2w'] > 5) | (jobs['log2h'] > 5))
    jobs['mts_idx'][big] = 0
    qpi = qp + 6 * (bit_depth - 8)
    jobs['qp_per'], jobs['qp_rem'] = qpi // 6, qpi % 6
    jobs['offset'] = offs
    jobs['lambda'] = lambda_for_qp(qp, bit_depth)
    resi = np.empty(total, np.int16)
    pred = np.empty(total, np.int16)
    # gather block by block, grouped by shape (vectorised per shape)
    for lw in range(2, 7):
        for lh in range(2, 7):
            idx = np.nonzero((jobs['log2w'] == lw) & (jobs['log2h'] == lh))[0]
            if not len(idx):
                continue
            w, h = 1 << lw, 1 << lh
            yy = jobs['y'][idx].astype(np.int64)[:, None, None] + np.arange(h)[None, :, None]
            xx = jobs['x'][idx].astype(np.int64)[:, None, None] + np.arange(w)[None, None, :]
            dst = offs[idx][:, None] + np.arange(w * h)[None, :]
            resi[dst] = resi_pic[yy, xx].reshape(len(idx), -1)
            pred[dst] = pred_pic[yy, xx].reshape(len(idx), -1)
    return jobs, resi, pred, default_dq_rates()


def slots_of_modes(visits, modes):
    """Evaluation slot (include/vvc_intra_b200.h VVCB_SLOT_*) of one candidate mode per visit.  modes: array of the
    result lists' vvcb_mode entries (fields mip, mrl, mode), one per visit."""
    from .engine import SLOT_MRL1, SLOT_MRL3, SLOT_MIP
    mip, mrl, mode = modes['mip'].astype(np.int64), modes['mrl'].astype(np.int64), modes['mode'].astype(np.int64)
    slot = mode.copy()
    slot[mip != 0] = SLOT_MIP + mode[mip != 0]
    has_mrl = (mip == 0) & (mrl != 0)
    if has_mrl.any():
        # MRL candidates are the visit's MPM[1..5] (EL/IntraSearch.cpp:635-681): the slot is indexed by the MPM position
        pos = (visits['mpm'][:, 1:].astype(np.int64) == mode[:, None]).argmax(axis=1)
        base = np.where(mrl == 1, SLOT_MRL1, SLOT_MRL3)
        slot[has_mrl] = (base + pos)[has_mrl]
    return slot.astype(np.uint8)


def default_ctx_states(seed=0):
    """A fixed pseudo-random estimator state (see default_dq_rates): every model somewhere between 'strongly 0' and 'strongly 1'."""
    from .engine import CTX_STATES_DTYPE, BIN_MODEL_DTYPE
    rng = np.random.default_rng(seed)
    st = np.zeros(1, CTX_STATES_DTYPE)
    flat = st.view(BIN_MODEL_DTYPE).reshape(1, -1)
    p = rng.integers(2000, 30000, flat.shape)
    flat['state'][..., 0] = p & 0x7fe0
    flat['state'][..., 1] = p & 0x7ffe
    flat['rate'] = (4 << 4) | 7                     # window sizes of the shipped initialisation tables lie around these
    return st


def build_tu_jobs_from_lists(visits, results, qp, bit_depth, rank=0, dep_quant=True, rate=False):
    """One DCT-II TU job per visit for the rank-th entry of its full-RD candidate list (results['final_mode']): what
    xRecurIntraCodingLumaQT hands to xIntraCodingTUBlock for that candidate.  Prediction and residual are left to the engine
    (vvcb_tu_eval_pred).  Returns (src, jobs, n_samples, rates)."""
    from .engine import TU_SRC_DTYPE
    n = len(visits)
    k = np.minimum(rank, results['n_final'] - 1)
    modes = results['final_mode'][np.arange(n), k]
    src = np.zeros(n, TU_SRC_DTYPE)
    src['visit'] = np.arange(n)
    src['slot'] = slots_of_modes(visits, modes)
    jobs = np.zeros(n, TU_JOB_DTYPE)
    for name in ('x', 'y', 'log2w', 'log2h'):
        jobs[name] = visits[name]
    jobs['flags'] = TU_QUANT | (TU_DEPQUANT if dep_quant else 0)
    if rate:                                            # TU::isTSAllowed / isMTSAllowed: sides <= 32 (CL/UnitTools.cpp:4524-4565)
        from .engine import TU_RATE, TU_TS_ALLOWED, TU_MTS_ALLOWED
        small = (visits['log2w'] <= 5) & (visits['log2h'] <= 5)
        jobs['flags'] |= TU_RATE
        jobs['flags'][small] |= TU_TS_ALLOWED | TU_MTS_ALLOWED
    qpi = qp + 6 * (bit_depth - 8)
    jobs['qp_per'], jobs['qp_rem'] = qpi // 6, qpi % 6
    sizes = (1 << visits['log2w'].astype(np.int64)) * (1 << visits['log2h'].astype(np.int64))
    jobs['offset'] = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    jobs['lambda'] = lambda_for_qp(qp, bit_depth)
    return src, jobs, int(sizes.sum()), default_dq_rates()
