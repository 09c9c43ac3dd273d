"""CPU tests: the dependent-quantisation oracle (oracle/vvc_oracle_dq.c) against the unmodified reference encoder's own
DepQuant::quant calls ('D' records: coefficients in, the estimator's context prices, levels and absSum out)."""
import collections

import numpy as np
import pytest

from oracle import oracle_py as O
import golden_util as G

FIXTURES = ['ref_10b_128x128_qp27_depquant', 'ref_8b_128x64_qp37_depquant']


@pytest.mark.parametrize('name', FIXTURES)
def test_dep_quant_levels_match_reference(name):
    _, tus = G.load_fixture(name)
    recs = [r for r in tus if r['tag'] == 'D']
    seen = collections.Counter()
    for r in recs:
        lvl, s = O.dep_quant(r['coeff'], r['bd'], r['mts'], r['lfnst'], r['qp'], r['lambda'], O.dq_rates_from_flat(r['rates']), r['cbf_delta'])
        assert s == r['abs_sum'], (r['w'], r['h'], r['mts'], s, r['abs_sum'])
        assert np.array_equal(lvl, r['level']), (r['w'], r['h'], r['mts'])
        seen[(r['w'], r['h'], r['mts'] > 1)] += 1
    assert len(seen) >= 20 and sum(1 for r in recs if r['abs_sum'] > 0) > 50
    assert any(k[0] == 64 for k in seen) and any(k[2] and 32 in k[:2] for k in seen)     # 64-point zero-out and MTS 32->16 zero-out


def test_dep_dequant_round_trip_properties():
    """Quantizer::dequantBlock: zero levels give zero; the reconstruction of the recorded levels stays within one quantiser step
    of the recorded coefficients wherever a level was kept (the trellis may zero a coefficient, never move it further)."""
    _, tus = G.load_fixture(FIXTURES[0])
    n = 0
    for r in [t for t in tus if t['tag'] == 'D' and t['abs_sum'] > 0][:60]:
        deq = O.dep_dequant(r['level'], r['bd'], r['qp'])
        assert not deq[r['level'] == 0].any()
        assert np.array_equal(np.sign(deq), np.sign(r['level']))
        nz = r['level'] != 0
        # the step of the two interleaved quantisers is 2 * Delta; |c - rec| <= 2 * Delta with Delta = rec / (2 |level| - 1 or so)
        step = np.abs(deq[nz]) / np.maximum(1, 2 * np.abs(r['level'][nz]) - 1)
        assert np.all(np.abs(r['coeff'][nz] - deq[nz]) <= 2.5 * step + 2)
        n += 1
    assert n > 20
    assert not O.dep_dequant(np.zeros((8, 8), np.int32), 10, 30).any()
