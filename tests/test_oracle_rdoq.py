"""CPU tests: the transform-skip RDOQ oracle (oracle/vvc_oracle_rdoq.c) against the unmodified reference encoder's own
QuantRDOQ::xRateDistOptQuantTS calls ('T' records: transform-skip coefficients and context prices in, levels and absSum out)."""
import numpy as np
import pytest

from oracle import oracle_py as O
import golden_util as G


@pytest.mark.parametrize('name', ['ref_10b_128x128_qp27_rdoqts', 'ref_8b_128x64_qp37_rdoqts'])
def test_rdoq_ts_levels_match_reference(name):
    _, tus = G.load_fixture(name)
    recs = [r for r in tus if r['tag'] == 'T']
    assert len(recs) > 100
    shapes = set()
    for r in recs:
        assert np.array_equal(O.fwd_transform(r['resi'], r['bd'], 1), r['coeff'])
        lvl, s = O.rdoq_ts(r['coeff'], r['bd'], r['qp'], r['lambda'], O.dq_rates_from_flat(None, r['rates']))
        assert s == r['abs_sum'], (r['w'], r['h'], s, r['abs_sum'])
        assert np.array_equal(lvl, r['level']), (r['w'], r['h'])
        shapes.add((r['w'], r['h']))
    assert len(shapes) >= 9
