"""CPU tests: the C-ABI shared library loads and exports every symbol include/vvc_intra_b200.h declares, the
ctypes mirror agrees with the header's struct sizes, and the product does not route through the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    g.build()
    import vvc_intra_b200 as vb
    return vb


def test_every_declared_symbol_is_exported(built):
    hdr = open(os.path.join(ROOT, 'include/vvc_intra_b200.h')).read()
    names = sorted(set(re.findall(r'\b(vvcb_[a-z_]+)\s*\(', hdr)))
    assert len(names) >= 16
    lib = C.CDLL(built.library_path())
    for n in names:
        assert hasattr(lib, n), n


def test_struct_sizes_match_header(built):
    src = '#include <stdio.h>\n#include "include/vvc_intra_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vvcb_rmd_visit), sizeof(vvcb_rmd_result), sizeof(vvcb_rates), sizeof(vvcb_cu_request), sizeof(vvcb_cu_auto), sizeof(vvcb_rmd_brief), sizeof(vvcb_rect));return 0;}'
    exe = os.path.join(ROOT, 'tests/host_emul/_sizes')
    subprocess.run(['gcc', '-x', 'c', '-', '-I', ROOT, '-o', exe], input=src.encode(), cwd=ROOT, check=True)
    out = subprocess.check_output([exe]).split()
    os.remove(exe)
    assert int(out[0]) == built.VISIT_DTYPE.itemsize == 80
    assert int(out[1]) == built.RESULT_DTYPE.itemsize == 368
    assert int(out[2]) == 44
    import ctypes
    from vvc_intra_b200 import engine as E
    assert int(out[3]) == ctypes.sizeof(E.CuRequest) and int(out[4]) == E.CU_AUTO_DTYPE.itemsize == 40
    assert int(out[5]) == E.BRIEF_DTYPE.itemsize == 64 and int(out[6]) == E.RECT_DTYPE.itemsize == 12


def test_no_cpu_fallback_without_device(built):
    """Without a CUDA device creating a context must fail loudly (skipped on a GPU box)."""
    lib = C.CDLL(built.library_path())
    if lib.vvcb_device_count() > 0:
        pytest.skip('a CUDA device is present')
    with pytest.raises(built.EngineError, match='no usable CUDA device'):
        built.IntraCostEngine(device=0)


def test_product_does_not_touch_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'vvc_intra_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle_py' not in txt and 'vvc_oracle' not in txt and 'libvvc_oracle' not in txt, f
    out = subprocess.check_output(['ldd', os.path.join(ROOT, 'vvc_intra_b200/libvvc_intra_b200.so')]).decode()
    assert 'oracle' not in out


def test_candidate_census(built):
    """1 345 reachable luma CUs per 64x64 root, 112 640 samples."""
    c = built.enumerate_root_candidates()
    assert len(c) == 1345 and int((c[:, 2] * c[:, 3]).sum()) == 112640
    census = {(4, 4): 256, (4, 8): 192, (8, 4): 192, (4, 16): 96, (16, 4): 96, (4, 32): 32, (32, 4): 32, (8, 8): 160,
              (8, 16): 84, (16, 8): 84, (8, 32): 28, (32, 8): 28, (16, 16): 36, (16, 32): 12, (32, 16): 12, (32, 32): 4,
              (64, 64): 1}
    got = {}
    for x, y, w, h in c:
        got[(int(w), int(h))] = got.get((int(w), int(h)), 0) + 1
    assert got == census


def test_sweep_visits_are_well_formed(built):
    v = built.build_sweep_visits(416, 240)
    w, h = 1 << v['log2w'].astype(int), 1 << v['log2h'].astype(int)
    assert np.all(v['x'] + w <= 416) and np.all(v['y'] + h <= 240)
    assert np.all(v['n_above'] <= w // 4) and np.all(v['n_left'] <= h // 4)
    # first CU of the picture sees nothing; a CU in the interior sees its left and above neighbours
    first = v[(v['x'] == 0) & (v['y'] == 0)]
    assert np.all(first['avail_al'] == 0) and np.all(first['n_above'] == 0) and np.all(first['n_left'] == 0)
    inner = v[(v['x'] == 64) & (v['y'] == 128) & (v['log2w'] == 6)]   # above-right lies in the previous CTU row
    assert inner['avail_al'][0] == 1 and inner['n_above'][0] == 16 and inner['n_left'][0] == 16
    assert inner['n_above_right'][0] == 16 and inner['n_below_left'][0] == 0


def test_calc_rd_cost_is_the_reference_formula(built):
    """vvcb_calc_rd_cost = RdCost::calcRdCost (CL/RdCost.cpp:63-74): (32768 / lambda) * dist + bits in IEEE double, that order (host logic)."""
    import random
    r = random.Random(7)
    for _ in range(2000):
        lam = r.uniform(0.5, 4000.0)
        bits, dist = r.randrange(0, 1 << 40), r.randrange(0, 1 << 36)
        scale = float(1 << 15) / lam
        assert built.IntraCostEngine.calc_rd_cost(lam, bits, dist) == scale * float(dist) + float(bits)
    assert built.IntraCostEngine.calc_rd_cost(57.0, 0, 0) == 0.0


def test_calc_rd_cost_matches_the_reference_function(built):
    """Bit patterns of RdCost::calcRdCost as the UNMODIFIED reference computes it (tests/golden/rd_cost.txt, written by
    oracle/dump_rd_cost.cpp through oracle/_ref/libvtmref.a: setLambda + saveUnadjustedLambda + calcRdCost, CL/RdCost.cpp:63-88)."""
    import struct
    n = 0
    for line in open(os.path.join(ROOT, 'tests/golden/rd_cost.txt')):
        if line.startswith('#'):
            continue
        lam_bits, bits, dist, cost_bits = line.split()
        lam = struct.unpack('<d', struct.pack('<Q', int(lam_bits, 16)))[0]
        got = built.IntraCostEngine.calc_rd_cost(lam, int(bits), int(dist))
        assert struct.unpack('<Q', struct.pack('<d', got))[0] == int(cost_bits, 16), line
        n += 1
    assert n == 600
