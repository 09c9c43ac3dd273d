"""CPU tests: TU-coding oracle (oracle/vvc_oracle_tr.c) against the reference encoder's own forward transforms,
MTS pre-selection, scalar quantiser and dequant + inverse transform (records 'S', 'Q', 'I' of the golden traces)."""
import collections

import numpy as np
import pytest

from oracle import oracle_py as O
import golden_util as G

ALL = ['ref_8b_128x64_qp32', 'ref_10b_192x128_qp27', 'ref_10b_64x64_qp32_scalarq']


@pytest.mark.parametrize('name', ALL)
def test_forward_transforms_and_mts_preselection(name):
    """a12: DCT-II / DST-VII / DCT-VIII / transform skip coefficients and the Sigma|coeff| candidate selection."""
    _, tus = G.load_fixture(name)
    seen = collections.Counter()
    for r in [t for t in tus if t['tag'] == 'S']:
        sums = []
        for m in r['modes']:
            got = O.fwd_transform(r['resi'], r['bd'], m['mts'])
            assert np.array_equal(got, m['coeff']), (r['w'], r['h'], m['mts'])
            sums.append(O.abs_sum_for_preselection(got, m['mts']))
            seen[(r['w'], r['h'], m['mts'])] += 1
        sel = O.mts_preselect(sums, r['w'], r['h'], r['max_cand'])
        assert sel == [m['selected'] for m in r['modes']], (r['w'], r['h'], sums)
    assert len(seen) >= 20


@pytest.mark.parametrize('name', ALL)
def test_forward_transform_of_quant_records(name):
    _, tus = G.load_fixture(name)
    n = 0
    for r in [t for t in tus if t['tag'] == 'Q']:
        got = O.fwd_transform(r['resi'], r['bd'], r['mts'])
        assert np.array_equal(got, r['coeff']), (r['w'], r['h'], r['mts'], r['load_tr'])
        n += 1
    assert n > 20


def test_scalar_quantiser():
    """a13 (scalar part): Quant::quant as run by the reference with DepQuant / RDOQ / sign hiding off."""
    _, tus = G.load_fixture('ref_10b_64x64_qp32_scalarq')
    n = 0
    for r in [t for t in tus if t['tag'] == 'Q']:
        assert r['dep_quant'] == 0
        lvl, s = O.quant_scalar(r['coeff'], r['bd'], r['per'], r['rem'], r['mts'] == 1)
        assert np.array_equal(lvl, r['level']), (r['w'], r['h'], r['mts'])
        assert s == r['abs_sum']
        n += 1
    assert n > 30


def test_dequant_and_inverse_transform():
    """a14: Quant::dequant + TrQuant::xIT / xITransformSkip.  Only the run without dependent quantisation applies:
    with DepQuant on, the reference reconstructs levels through DepQuant::dequant's state machine instead."""
    _, tus = G.load_fixture('ref_10b_64x64_qp32_scalarq')
    n = 0
    for r in [t for t in tus if t['tag'] == 'I']:
        co = O.dequant(r['level'], r['bd'], r['per'], r['rem'], r['mts'] == 1)
        resi = O.inv_transform(co, r['bd'], r['mts'])
        assert np.array_equal(resi, r['resi']), (r['w'], r['h'], r['mts'])
        n += 1
    assert n > 20
