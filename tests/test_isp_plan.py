"""Intra sub-partition geometry (vvcb_isp_plan, pure host logic of the library) against the reference's own functions.

tests/golden/isp_geometry.txt was written by oracle/dump_isp_geometry.cpp (`make -f oracle/Makefile.ref isp_geometry`), which calls
the UNMODIFIED reference's CU::canUseISP, CU::getISPSplitDim, CU::isMinWidthPredEnabledForBlkSize and TrQuant::getTrTypes
(CL/UnitTools.cpp:426-460, :4342; CL/TrQuant.cpp:752) out of oracle/_ref/libvtmref.a for every CU size x split x maximum
transform size.  No GPU is involved: the planner is host code of libvvc_intra_b200.so."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests/golden/isp_geometry.txt')


@pytest.fixture(scope='module')
def eng():
    import __graft_entry__ as g
    g.build()
    import vvc_intra_b200 as V
    return V


def _cases():
    for line in open(GOLDEN):
        if line.startswith('#') or not line.strip():
            continue
        head, _, tail = line.partition('|')
        h = [int(v) for v in head.split()]
        t = [int(v) for v in tail.split()]
        yield h, [t[i:i + 8] for i in range(0, len(t), 8)]


def test_golden_covers_every_size_split_and_transform_limit():
    cases = list(_cases())
    assert len(cases) == 5 * 5 * 2 * 2
    assert sum(1 for h, _ in cases if h[4]) == 2 * (15 + 24)      # 4x4 never; sides of 64 only with max_tb 64: 15 and 24 sizes, two splits each


def test_plan_matches_the_reference_functions(eng):
    E = eng.IntraCostEngine
    checked = 0
    for (w, h, max_tb, split, allowed, *rest), blocks in _cases():
        for use_mts in (True, False):
            parts = E.isp_plan(w, h, split, max_tb, use_mts)
            if not allowed:
                assert len(parts) == 0, (w, h, max_tb, split)
                continue
            size, n, min_width_pred = rest
            assert len(parts) == n == len(blocks)
            for p, (x, y, bw, bh, th1, tv1, th0, tv0) in zip(parts, blocks):
                assert (p['x'], p['y'], p['w'], p['h']) == (x, y, bw, bh)
                assert (p['tr_hor'], p['tr_ver']) == ((th1, tv1) if use_mts else (th0, tv0))
                # CU::isPredRegDiffFromTB / adjustPredArea / isFirstTBInPredReg (CL/UnitTools.cpp:4334-4355)
                if min_width_pred:
                    assert p['pred_w'] == max(4, bw) and p['pred_x'] == x - x % 4 and p['predicts'] == (x % 4 == 0)
                else:
                    assert (p['pred_x'], p['pred_w'], p['predicts']) == (x, bw, 1)
                assert (p['pred_y'], p['pred_h']) == (y, bh)
                # initIntraPatternChTypeISP (CL/IntraPrediction.cpp:1132-1133, :1142-1143)
                assert p['top_ref_len'] == w + p['pred_w'] and p['left_ref_len'] == h + p['pred_h']
                checked += 1
            # the one fetch over the CU (CL/IntraPrediction.cpp:1115-1124)
            f = parts[0]
            if split == eng.ISP_HOR:
                assert (f['fetch_top_len'], f['fetch_left_len']) == (w + f['pred_w'], 2 * h)
            else:
                assert (f['fetch_top_len'], f['fetch_left_len']) == (2 * w, h + f['pred_h'])
            assert not parts[1:]['fetch_top_len'].any() and not parts[1:]['fetch_left_len'].any()
            assert list(parts['last']) == [0] * (n - 1) + [1]
    assert checked > 400


def test_plan_properties(eng):
    """Size-independent properties: the blocks tile the CU, each holds at least 16 samples, prediction regions tile it too."""
    E = eng.IntraCostEngine
    for w in (4, 8, 16, 32, 64):
        for h in (4, 8, 16, 32, 64):
            for split in (eng.ISP_HOR, eng.ISP_VER):
                parts = E.isp_plan(w, h, split)
                if w * h <= 16:
                    assert len(parts) == 0
                    continue
                assert len(parts) in (2, 4)
                cover = np.zeros((h, w), np.int32)
                pred = np.zeros((h, w), np.int32)
                for p in parts:
                    assert p['w'] * p['h'] >= 16
                    cover[p['y']:p['y'] + p['h'], p['x']:p['x'] + p['w']] += 1
                    if p['predicts']:
                        assert p['pred_w'] >= 4 or split == eng.ISP_HOR or p['pred_w'] == p['w']
                        pred[p['pred_y']:p['pred_y'] + p['pred_h'], p['pred_x']:p['pred_x'] + p['pred_w']] += 1
                assert (cover == 1).all() and (pred == 1).all()


def test_bad_arguments_are_refused(eng):
    E = eng.IntraCostEngine
    for args in ((12, 8, 1), (8, 128, 1), (1 << 30, 8, 1), (8, 2147483647, 2), (-8, 8, 1), (2, 8, 2), (8, 8, 0), (8, 8, 3)):
        with pytest.raises(eng.EngineError):
            E.isp_plan(*args)


def test_struct_size_matches_header(eng):
    src = '#include <stdio.h>\n#include "include/vvc_intra_b200.h"\nint main(){printf("%zu\\n", sizeof(vvcb_isp_part));return 0;}'
    exe = os.path.join(ROOT, 'tests/host_emul/_isp_size')
    subprocess.run(['gcc', '-x', 'c', '-', '-I', ROOT, '-o', exe], input=src.encode(), cwd=ROOT, check=True)
    out = int(subprocess.check_output([exe]))
    os.remove(exe)
    assert out == eng.ISP_PART_DTYPE.itemsize == 28


def test_isp_mode_params_match_the_reference_function(eng):
    """vvcb_isp_mode_param against the UNMODIFIED reference's compiled IntraPrediction::initPredIntraParams with cu.ispMode set, for the
    prediction regions of every ISP CU shape x both splits x 67 modes (the ISP part of tests/golden/intra_params.txt.gz, written by
    oracle/dump_intra_params.cpp).  The region sizes of the table are the ones vvcb_isp_plan returns."""
    import gzip
    E = eng.IntraCostEngine
    n = 0
    regions = {}
    for line in gzip.open(os.path.join(ROOT, 'tests/golden/intra_params.txt.gz'), 'rt'):
        if not line.startswith('ISP'):
            continue
        head, _, tail = line[3:].partition('|')
        w, h, split, pw, ph, mode = (int(v) for v in head.split())
        is_ver, ref_filter, interp, pdpc, angle, inv_angle, scale = (int(v) for v in tail.split())
        if (w, h, split) not in regions:
            regions[(w, h, split)] = E.isp_plan(w, h, split)[0]
        first = regions[(w, h, split)]
        assert (first['pred_w'], first['pred_h']) == (pw, ph)
        p = E.isp_mode_param(w, h, pw, ph, mode)
        key = (w, h, split, pw, ph, mode)
        assert (ref_filter, interp) == (0, 0), key          # never smoothed, cubic interpolation: what the header promises
        assert (p['is_ver'], p['pdpc']) == (is_ver, pdpc), key
        if mode > 1:
            assert (p['angle'], p['inv_angle']) == (angle, inv_angle), key
            if angle > 0 and pdpc:
                assert p['ang_scale'] == scale, key
        else:
            assert p['angle'] == 0
        n += 1
    assert n == 48 * 67 and len(regions) == 48
    for args in ((8, 8, 16, 2, 5), (8, 8, 8, 2, 67), (6, 8, 2, 8, 3), (128, 8, 8, 8, 3)):
        with pytest.raises(eng.EngineError):
            E.isp_mode_param(*args)
