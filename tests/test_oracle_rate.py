"""CPU tests: the residual rate-estimation oracle (oracle/vvc_oracle_rate.c) against the unmodified reference encoder's own
CABACWriter::residual_coding calls on its bit estimator ('C' records: levels and the estimator's context states in, the fractional
bits the call added out)."""
import collections

import numpy as np
import pytest

from oracle import oracle_py as O
import golden_util as G


@pytest.mark.parametrize('name', ['ref_10b_128x128_qp27_resbits', 'ref_8b_128x64_qp22_resbits'])
def test_residual_bits_match_reference(name):
    _, tus = G.load_fixture(name)
    recs = [r for r in tus if r['tag'] == 'C']
    assert len(recs) > 150
    kinds = collections.Counter()
    for r in recs:
        got = O.residual_bits(r['level'], r['mts'], r['ts_allowed'], r['mts_allowed'], r['dep_quant'], O.ctx_states_from_record(r['states']))
        assert got == r['bits'], (r['w'], r['h'], r['mts'], got, r['bits'])
        kinds[(r['w'], r['h'], min(r['mts'], 2))] += 1
    assert len(kinds) >= 30 and any(k[2] == 1 for k in kinds) and any(k[2] == 2 for k in kinds)
