"""Host logic of the exhaustive candidate sweep: which luma CUs can occur under a 64x64 search root with
the reference's shipped configuration, and what reference-sample availability each one sees.

Restates QTBTPartitioner::canSplit (CL/UnitPartitioner.cpp:379-466) for the all-intra luma tree of
BIN/encoder_intra.cfg: CTU 128, dual tree => 128x128 is force-split to 64x64 (EL/EncModeCtrl.cpp:1647),
MinQTLumaISlice 8, MaxBT = MaxTT = 32 (CL/CommonDef.h:427-429), MaxMTTHierarchyDepth 3, min CU side 4.
SURVEY.md App. C gives the expected census: 1 345 distinct areas per root.
"""
import functools

import numpy as np

from .engine import VISIT_DTYPE

MIN_QT, MAX_BT, MAX_TT, MAX_MTT_DEPTH, MIN_SIDE = 8, 32, 32, 3, 4


@functools.lru_cache(maxsize=None)
def enumerate_root_candidates(root=64):
    """All distinct (x, y, w, h) reachable below one root, relative to the root origin, sorted."""
    seen = set()

    def visit(x, y, w, h, mtt_depth, qt_ok, no_bt_h, no_bt_v):
        if w <= 64 and h <= 64:
            seen.add((x, y, w, h))
        # quad split: only while no MTT split happened above, square, larger than MinQT
        if qt_ok and w == h and w > MIN_QT:
            hw = w // 2
            for dy in (0, hw):
                for dx in (0, hw):
                    visit(x + dx, y + dy, hw, hw, 0, True, False, False)
        if mtt_depth >= MAX_MTT_DEPTH:
            return
        if w <= MAX_BT and h <= MAX_BT:
            if h > MIN_SIDE and not no_bt_h:            # BT horizontal
                for k in range(2):
                    visit(x, y + k * h // 2, w, h // 2, mtt_depth + 1, False, False, False)
            if w > MIN_SIDE and not no_bt_v:            # BT vertical
                for k in range(2):
                    visit(x + k * w // 2, y, w // 2, h, mtt_depth + 1, False, False, False)
        if w <= MAX_TT and h <= MAX_TT:
            if h >= 4 * MIN_SIDE:                       # TT horizontal: h/4, h/2, h/4
                q = h // 4
                visit(x, y, w, q, mtt_depth + 1, False, False, False)
                visit(x, y + q, w, 2 * q, mtt_depth + 1, False, True, False)   # middle part: no BT in the same direction
                visit(x, y + 3 * q, w, q, mtt_depth + 1, False, False, False)
            if w >= 4 * MIN_SIDE:                       # TT vertical
                q = w // 4
                visit(x, y, q, h, mtt_depth + 1, False, False, False)
                visit(x + q, y, 2 * q, h, mtt_depth + 1, False, False, True)
                visit(x + 3 * q, y, q, h, mtt_depth + 1, False, False, False)

    visit(0, 0, root, root, 0, True, False, False)
    arr = np.array(sorted(seen, key=lambda c: (-c[2] * c[3], c[1], c[0], c[2])), dtype=np.int32)
    return arr


def frame_candidates(width, height, root=64):
    """Candidates of every root of a picture that lie completely inside it: int32 array [n, 4]."""
    base = enumerate_root_candidates(root)
    out = []
    for ry in range(0, height, root):
        for rx in range(0, width, root):
            c = base.copy()
            c[:, 0] += rx
            c[:, 1] += ry
            keep = (c[:, 0] + c[:, 2] <= width) & (c[:, 1] + c[:, 3] <= height)
            out.append(c[keep])
    return np.concatenate(out)


def _morton(x, y):
    """z-order index of the 4x4 block at sample position (x, y) inside its CTU (x, y < 128)."""
    x = (x >> 2).astype(np.int64)
    y = (y >> 2).astype(np.int64)
    r = np.zeros_like(x)
    for b in range(5):
        r |= ((x >> b) & 1) << (2 * b)
        r |= ((y >> b) & 1) << (2 * b + 1)
    return r


def candidate_availability(cands, width, height, ctu=128):
    """Reference-sample availability of each candidate, in 4-sample units, under the decoding order of
    the reference: CTUs in raster order (EL/EncSlice.cpp:1608), z-order inside a CTU.  A neighbouring unit
    is available iff it lies inside the picture and precedes the candidate in that order -- the prefix
    counts that isAboveAvailable & co. (CL/IntraPrediction.cpp:1524-1662) would return for a quad-tree
    shaped neighbourhood.  Returns dict of uint8 arrays."""
    x, y, w, h = (cands[:, i].astype(np.int64) for i in range(4))
    n = len(cands)

    def before(px, py):
        """is sample (px, py) decoded before the candidate at (x, y)?"""
        inside = (px >= 0) & (py >= 0) & (px < width) & (py < height)
        pcx, pcy = np.floor_divide(px, ctu), np.floor_divide(py, ctu)
        ccx, ccy = x // ctu, y // ctu
        earlier_ctu = (pcy < ccy) | ((pcy == ccy) & (pcx < ccx))
        same_ctu = (pcy == ccy) & (pcx == ccx)
        z_p = _morton(np.where(inside, px % ctu, 0), np.where(inside, py % ctu, 0))
        z_c = _morton(x % ctu, y % ctu)
        return inside & (earlier_ctu | (same_ctu & (z_p < z_c)))

    def prefix(count_max, pos_of):
        cnt = np.zeros(n, np.int64)
        alive = np.ones(n, bool)
        for u in range(16):
            px, py = pos_of(u)
            ok = alive & (u < count_max) & before(px, py)
            cnt += ok
            alive = ok
        return cnt.astype(np.uint8)

    return dict(
        avail_al=before(x - 1, y - 1).astype(np.uint8),
        n_above=prefix(w // 4, lambda u: (x + 4 * u, y - 1)),
        n_above_right=prefix(w // 4, lambda u: (x + w + 4 * u, y - 1)),
        n_left=prefix(h // 4, lambda u: (x - 1, y + 4 * u)),
        n_below_left=prefix(h // 4, lambda u: (x - 1, y + h + 4 * u)),
    )


# A context snapshot recorded from the reference at QP 32 (tests/golden/ref_8b_128x64_qp32, first visit):
# mip_flag[2], mrl_bin0[2], mrl_bin1[2], isp_bin0_0, mpm_flag[2], planar_flag[2]
DEFAULT_RATES = (7705, 89562, 4349, 114955, 38452, 27695, 7705, 70599, 12026, 33700, 31854)
DEFAULT_MPM = (0, 1, 50, 18, 46, 54)      # PU::getIntraMPMs with planar neighbours


def sqrt_lambda_for_qp(qp):
    """RdCost::getMotionLambda() * FRAC_BITS_SCALE as the reference prints it for an intra picture at QP 32
    (0.00023903622867981877, recorded), scaled by 2^((qp-32)/6) like sqrt(lambda)."""
    return 0.00023903622867981877 * 2.0 ** ((qp - 32) / 6.0)


def build_sweep_visits(width, height, qp=32, ctu=128, rates=DEFAULT_RATES, mpm=DEFAULT_MPM):
    """Visit descriptors of the exhaustive sweep of one picture: every candidate CU of every 64x64 root,
    big CUs first inside each root."""
    c = frame_candidates(width, height)
    av = candidate_availability(c, width, height, ctu)
    v = np.zeros(len(c), VISIT_DTYPE)
    v['x'], v['y'] = c[:, 0], c[:, 1]
    v['log2w'] = np.log2(c[:, 2]).astype(np.uint8)
    v['log2h'] = np.log2(c[:, 3]).astype(np.uint8)
    for k, a in av.items():
        v[k] = a
    v['mpm'] = np.array(mpm, np.uint8)
    v['num_mpm_cand'] = 1
    v['rates'] = np.array(rates, np.uint32)
    v['sqrt_lambda'] = sqrt_lambda_for_qp(qp)
    return v
