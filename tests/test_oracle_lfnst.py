"""CPU tests: the LFNST oracle (oracle/vvc_oracle_tr.c) against the unmodified reference encoder ('F' records: residual in,
coefficients after primary transform + forward LFNST and dependent-quantisation levels out; 'J' records: levels in, residual after
dependent dequantisation + inverse LFNST + inverse primary transform out)."""
import numpy as np
import pytest

from oracle import oracle_py as O
import golden_util as G

FIXTURES = ['ref_10b_128x128_qp27_lfnst', 'ref_8b_128x64_qp32_lfnst']


@pytest.mark.parametrize('name', FIXTURES)
def test_forward_lfnst_and_quantisation(name):
    _, tus = G.load_fixture(name)
    recs = [r for r in tus if r['tag'] == 'F']
    assert len(recs) > 50
    kinds = set()
    for r in recs:
        co = O.fwd_lfnst(O.fwd_transform(r['resi'], r['bd'], r['mts'], r['lfnst']), r['intra_mode'], r['lfnst'])
        # outside the top-left 8x8 (4x4) the reference's buffer keeps stale values of earlier transforms for some sizes: the
        # partial butterflies skip those outputs (CL/TrQuant.cpp:853-867) and nothing downstream reads them (the quantiser stops
        # at scan position 7 / 15, CL/DepQuant.cpp:1641-1646); the restatement writes zeros there
        sb = 8 if min(r['w'], r['h']) >= 8 else 4
        assert np.array_equal(co[:sb, :sb], r['coeff'][:sb, :sb]), (r['w'], r['h'], r['intra_mode'], r['lfnst'])
        assert not co[sb:, :].any() and not co[:, sb:].any()
        lvl, s = O.dep_quant(co, r['bd'], r['mts'], r['lfnst'], r['qp'], r['lambda'], O.dq_rates_from_flat(r['rates']), r['cbf_delta'])
        assert s == r['abs_sum'] and np.array_equal(lvl, r['level']), (r['w'], r['h'], r['intra_mode'], r['lfnst'])
        kinds.add((min(r['w'], 8), min(r['h'], 8), r['lfnst'], r['intra_mode'] > 34))
    assert len(kinds) >= 10


@pytest.mark.parametrize('name', FIXTURES)
def test_inverse_lfnst(name):
    _, tus = G.load_fixture(name)
    recs = [r for r in tus if r['tag'] == 'J']
    assert len(recs) > 50
    for r in recs:
        co = O.inv_lfnst(O.dep_dequant(r['level'], r['bd'], r['qp']), r['intra_mode'], r['lfnst'])
        resi = O.inv_transform(co, r['bd'], r['mts'])
        assert np.array_equal(resi, r['resi']), (r['w'], r['h'], r['intra_mode'], r['lfnst'])
