"""Frame-parallel gather (vvc_intra_b200/assemble.py, vvc_intra_b200/hls.py): pictures encoded by independent encoder processes and gathered into the
sequential encoder's bitstream, byte for byte; the reference's own Parcat reproduced byte for byte.  The fixtures under tests/golden/assemble/ are
outputs of the unmodified reference (tools/make_assemble_golden.py); the live tests run the reference binaries built by oracle/Makefile.ref and are
skipped where those are absent."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle/_ref')
GOLD = os.path.join(ROOT, 'tests/golden/assemble')


def test_bit_exact_gather_reproduces_the_sequential_encoder(tmp_path):
    """Four 256x128 10-bit pictures encoded one process per picture; the gather must be the sequential encoder's stream.  The slice headers of
    pictures 1.. grow by two bits (one byte after re-alignment for pictures 1-3), picture 3 carries an ALF APS whose id must not move."""
    from vvc_intra_b200 import assemble, hls
    paths = [os.path.join(GOLD, 'pic_256x128_10b_qp27_f%d.bin' % f) for f in range(4)]
    stats = assemble.assemble_sequential(paths, str(tmp_path / 'all.bin'))
    seq = open(os.path.join(GOLD, 'pic_256x128_10b_qp27_seq.bin'), 'rb').read()
    assert (tmp_path / 'all.bin').read_bytes() == seq
    assert [s['bytes_out'] - s['bytes_in'] for s in stats] == [0, 1, 1, 1]
    types = [hls.NAL_NAMES[hls.nal_unit_type(u)] for _, u in hls.split_nal_units(seq)]
    assert types.count('IDR_W_RADL') == 1 and types.count('CRA') == 3 and types.count('APS') == 1
    # the plain concatenation differs from it in exactly the three re-numbered slice NAL units
    assemble.concat_segments(paths, str(tmp_path / 'cat.bin'))
    diff = assemble.diff_against_sequential((tmp_path / 'cat.bin').read_bytes(), seq)
    assert [(d[1], d[2]) for d in diff] == [('IDR_W_RADL', 'CRA')] * 3
    # ReWriteParamSets=0 flavour: parameter sets only ahead of the first picture
    assemble.assemble_sequential(paths, str(tmp_path / 'lean.bin'), rewrite_param_sets=False)
    lean = [hls.NAL_NAMES[hls.nal_unit_type(u)] for _, u in hls.split_nal_units((tmp_path / 'lean.bin').read_bytes())]
    assert lean.count('SPS') == 1 and lean.count('PPS') == 1 and lean.count('CRA') == 3
    # a segment that is not a one-picture IDR stream is refused
    with pytest.raises(ValueError):
        assemble.assemble_sequential([paths[0], os.path.join(GOLD, 'pic_256x128_10b_qp27_seq.bin')], str(tmp_path / 'x.bin'))


def test_bit_exact_gather_when_every_picture_carries_an_alf_aps(tmp_path):
    """configs[0]'s picture size and QP (416x240 8-bit QP 32), two pictures, each with an ALF APS: the sequential encoder gives both id 7 -- the id counter
    restarts with every CRA picture (pending-RAS initialisation, EL/EncGOP.cpp:4213-4225 -> EL/EncAdaptiveLoopFilter.cpp:667-674) -- so the per-picture
    streams already carry the right ids and the gather is byte-identical."""
    from vvc_intra_b200 import assemble, hls
    paths = [os.path.join(GOLD, 'pic_416x240_8b_qp32_f%d.bin' % f) for f in range(2)]
    assemble.assemble_sequential(paths, str(tmp_path / 'all.bin'))
    seq = open(os.path.join(GOLD, 'pic_416x240_8b_qp32_seq.bin'), 'rb').read()
    assert (tmp_path / 'all.bin').read_bytes() == seq
    aps = [hls.BitReader(hls.unescape(u[2:])) for _, u in hls.split_nal_units(seq) if hls.nal_unit_type(u) == hls.NAL_APS]
    assert [(r.u(5), r.u(3)) for r in aps] == [(7, 0), (7, 0)]            # adaptation_parameter_set_id, aps_params_type (ALF), EL/VLCWriter.cpp:493-513


def test_library_gather_matches_the_reference_and_its_python_twin(tmp_path):
    """include/vvc_intra_b200_gather.h: the C++ gather inside libvvc_intra_b200.so (host code; runs without a GPU) on the reference fixtures."""
    import ctypes as C
    import re
    from vvc_intra_b200 import assemble, engine
    lib = engine.load_library()
    hdr = open(os.path.join(ROOT, 'include/vvc_intra_b200_gather.h')).read()
    names = sorted(set(re.findall(r'\b(vvcb_gather_[a-z_]+)\s*\(', hdr)))
    assert names == ['vvcb_gather_parcat', 'vvcb_gather_sequential'] and all(hasattr(lib, n) for n in names)
    for stem, n in (('pic_256x128_10b_qp27', 4), ('pic_416x240_8b_qp32', 2)):
        paths = [os.path.join(GOLD, '%s_f%d.bin' % (stem, f)) for f in range(n)]
        seq = open(os.path.join(GOLD, stem + '_seq.bin'), 'rb').read()
        assert assemble.gather_sequential(paths, str(tmp_path / 'lib.bin')) == len(seq)
        assert (tmp_path / 'lib.bin').read_bytes() == seq
        assemble.gather_sequential(paths, str(tmp_path / 'lib_lean.bin'), rewrite_param_sets=False)
        assemble.assemble_sequential(paths, str(tmp_path / 'py_lean.bin'), rewrite_param_sets=False)
        assert (tmp_path / 'lib_lean.bin').read_bytes() == (tmp_path / 'py_lean.bin').read_bytes()
    segs = [os.path.join(GOLD, 'seg_64x64_8b_qp32_s%d.bin' % k) for k in range(3)]
    assert assemble.gather_parcat(segs, str(tmp_path / 'pc.bin')) == 6
    assert (tmp_path / 'pc.bin').read_bytes() == open(os.path.join(GOLD, 'seg_64x64_8b_qp32_parcat.bin'), 'rb').read()
    # error behaviour: status codes and messages of the C ABI
    with pytest.raises(ValueError, match='not a one-picture segment'):
        assemble.gather_sequential([os.path.join(GOLD, 'pic_256x128_10b_qp27_seq.bin')], str(tmp_path / 'x.bin'))
    with pytest.raises(ValueError, match='cannot open'):
        assemble.gather_sequential([str(tmp_path / 'missing.bin')], str(tmp_path / 'x.bin'))
    with pytest.raises(ValueError, match='cannot write'):
        assemble.gather_parcat(segs, str(tmp_path / 'no_such_dir' / 'x.bin'))
    (tmp_path / 'junk.bin').write_bytes(b'\x00\x00\x01\x91\x02' + bytes(20) + b'\x80')      # an SPS NAL unit of zeros
    with pytest.raises(ValueError):
        assemble.gather_parcat([str(tmp_path / 'junk.bin')], str(tmp_path / 'x.bin'))
    err = C.create_string_buffer(64)
    assert lib.vvcb_gather_sequential(None, 0, None, 1, None, err, 64) == -1 and b'bad argument' in err.value
    assert lib.vvcb_gather_parcat(None, 0, None, None, None, 0) == -1                               # no message buffer: still a status


def test_library_gather_survives_damaged_streams(tmp_path):
    """Mutation fuzz of the C++ gather under AddressSanitizer + UBSan (tests/host_emul/gather_fuzz.cpp): bit flips, replaced bytes, truncations and
    inserted bytes in reference streams must come back as status codes."""
    exe = str(tmp_path / 'gather_fuzz')
    r = subprocess.run(['g++', '-std=c++17', '-O1', '-g', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined', '-o', exe,
                        os.path.join(ROOT, 'tests/host_emul/gather_fuzz.cpp')], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        pytest.skip('no sanitizer runtime for g++ here: ' + r.stdout[-300:])
    streams = [os.path.join(GOLD, f) for f in ('pic_416x240_8b_qp32_f0.bin', 'pic_416x240_8b_qp32_f1.bin', 'seg_64x64_8b_qp32_s0.bin', 'seg_64x64_8b_qp32_s1.bin')]
    r = subprocess.run([exe, '2500', str(tmp_path)] + streams, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    ok, refused, unsupported = (int(v) for v in r.stdout.split())
    assert ok + refused + unsupported == 5000 and ok > 500 and refused > 500


def test_parcat_segments_matches_the_reference_tool(tmp_path):
    from vvc_intra_b200 import assemble
    segs = [os.path.join(GOLD, 'seg_64x64_8b_qp32_s%d.bin' % k) for k in range(3)]
    assert assemble.parcat_segments(segs, str(tmp_path / 'out.bin')) == 6                  # six pictures re-numbered (POC 1..6)
    assert (tmp_path / 'out.bin').read_bytes() == open(os.path.join(GOLD, 'seg_64x64_8b_qp32_parcat.bin'), 'rb').read()
    assert assemble.parcat_segments(segs[:1], str(tmp_path / 'one.bin')) == 2              # a single segment passes through unchanged
    assert (tmp_path / 'one.bin').read_bytes() == open(segs[0], 'rb').read()


def test_syntax_reader_on_reference_streams():
    """Every parameter set of the fixtures parses to its rbsp_trailing_bits, every slice header to its alignment bits; escaping is the encoder's."""
    import random
    from vvc_intra_b200 import hls
    seen = set()
    for name in sorted(os.listdir(GOLD)):
        data = open(os.path.join(GOLD, name), 'rb').read()
        sps, pps = {}, {}
        for _, u in hls.split_nal_units(data):
            t, rbsp = hls.nal_unit_type(u), hls.unescape(u[2:])
            assert hls.escape(rbsp) == u[2:]
            if t == hls.NAL_SPS:
                s = hls.parse_sps(rbsp)
                sps[s['sps_id']] = s
                assert (s['width'], s['height'], s['poc_bits']) in ((256, 128, 8), (64, 64, 8), (416, 240, 8)) and s['alf'] and s['sao'] and s['lmcs']
            elif t == hls.NAL_PPS:
                p = hls.parse_pps(rbsp, sps)
                pps[p['pps_id']] = p
            elif t in (hls.NAL_IDR_W_RADL, hls.NAL_CRA):
                h = hls.parse_intra_slice_header(rbsp, t, sps[0], pps)
                assert h['slice_type'] == hls.I_SLICE and (h['after_rpl'] - h['after_poc']) == (2 if t == hls.NAL_CRA else 0)
                seen.add((t, h['poc_lsb']))
    assert {(hls.NAL_CRA, n) for n in range(1, 7)} <= seen and (hls.NAL_IDR_W_RADL, 0) in seen
    rng = random.Random(7)
    for _ in range(200):                                             # emulation prevention round trip on zero-heavy payloads
        rbsp = bytes(rng.choice((0, 0, 0, 1, 2, 3, 4, 255)) for _ in range(rng.randrange(1, 40)))
        esc = hls.escape(rbsp)
        assert b'\x00\x00\x00' not in esc and b'\x00\x00\x01' not in esc and b'\x00\x00\x02' not in esc and esc[-1] != 0
        assert hls.unescape(esc) == rbsp or (rbsp[-1] == 0 and hls.unescape(esc) == rbsp + b'\x03')
    with pytest.raises(ValueError):
        hls.parse_sps(bytes(40))
    w = hls.BitWriter()
    for v in (0, 1, 2, 7, 8, 300):
        w.ue(v)
    w.align()
    r = hls.BitReader(w.tobytes())
    assert [r.ue() for _ in range(6)] == [0, 1, 2, 7, 8, 300] and r.at_trailing_bits()


def test_frame_parallel_gather_decodes_to_the_sequential_reconstruction(tmp_path):
    from make_golden import synth_yuv
    from vvc_intra_b200 import assemble
    enc, dec, cfg = (os.path.join(REF, f) for f in ('EncoderApp', 'DecoderApp', 'encoder_intra.cfg'))
    if not all(os.path.exists(p) for p in (enc, dec, cfg)):
        pytest.skip('oracle/_ref binaries are built only in the container that has /root/reference')
    w, h, bits, qp, n = 64, 64, 8, 32, 3
    data = b''
    for f in range(n):
        Y, U, V = synth_yuv(w, h, bits, f)
        data += Y.tobytes() + U.tobytes() + V.tobytes()
    (tmp_path / 'in.yuv').write_bytes(data)
    (tmp_path / 'Time_python.dat').write_bytes(b'')
    args = [enc, '-c', cfg, '-i', 'in.yuv', '-wdt', str(w), '-hgt', str(h), '-q', str(qp), '-fr', '30', '--InputBitDepth=%d' % bits,
            '--InternalBitDepth=%d' % bits, '--OutputBitDepth=%d' % bits]
    run = lambda cmd: subprocess.run(cmd, cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    # one encoder process per picture (they are independent: run concurrently), and the sequential encoder
    procs = [subprocess.Popen(args + ['-f', '1', '--FrameSkip=%d' % f, '-b', 'f%d.bin' % f, '-o', 'f%d.yuv' % f], cwd=tmp_path, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
             for f in range(n)]
    procs.append(subprocess.Popen(args + ['-f', str(n), '-b', 'seq.bin', '-o', 'seq.yuv'], cwd=tmp_path, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    assert all(p.wait() == 0 for p in procs)
    stats = assemble.concat_segments([str(tmp_path / ('f%d.bin' % f)) for f in range(n)], str(tmp_path / 'all.bin'))
    assert len(stats) == n and all('IDR_W_RADL' in s['nal_units'] or 'IDR_N_LP' in s['nal_units'] for s in stats)
    run([dec, '-b', 'all.bin', '-o', 'dec.yuv', '-d', str(bits)])
    decoded = (tmp_path / 'dec.yuv').read_bytes()
    per_picture = b''.join((tmp_path / ('f%d.yuv' % f)).read_bytes() for f in range(n))
    assert len(decoded) == n * w * h * 3 // 2 and decoded == per_picture                 # the gather decodes to what the encoders reconstructed
    assert decoded == (tmp_path / 'seq.yuv').read_bytes()                                  # ... which is the sequential encoder's reconstruction
    # what separates the gather from the sequential bitstream: only the slice NAL units of pictures 1.. (IDR / POC 0 against CRA / POC n), same sizes
    diff = assemble.diff_against_sequential((tmp_path / 'all.bin').read_bytes(), (tmp_path / 'seq.bin').read_bytes())
    assert len(diff) == n - 1 and all(d[1].startswith('IDR') and d[2] == 'CRA' and d[3] == d[4] for d in diff), diff
    # the bit-exact gather closes that gap
    assemble.assemble_sequential([str(tmp_path / ('f%d.bin' % f)) for f in range(n)], str(tmp_path / 'exact.bin'))
    assert (tmp_path / 'exact.bin').read_bytes() == (tmp_path / 'seq.bin').read_bytes()
    with pytest.raises(ValueError):
        (tmp_path / 'bad.bin').write_bytes(b'\x00\x00\x01\x11\x02')
        assemble.concat_segments([str(tmp_path / 'bad.bin')], str(tmp_path / 'x.bin'))


@pytest.mark.parametrize('w,h,bits,qp,extra', [
    (64, 64, 10, 22, ['--SEIDecodedPictureHash=1']),                 # suffix SEI behind every picture
    (64, 64, 8, 37, ['--ALF=0', '--SAO=0']),                         # shorter slice headers
    (128, 64, 8, 27, ['--LMCSEnable=0', '--DepQuant=0']),            # sign_data_hiding_enabled_flag in the header
    (64, 64, 8, 32, ['--LoopFilterDisable=1', '--JointCbCr=0']),     # deblocking control in the PPS
    (72, 40, 8, 32, []),                                             # conformance window
])
def test_bit_exact_gather_across_encoder_options(w, h, bits, qp, extra, tmp_path):
    """Live: the header reader must follow whatever the options do to the SPS / PPS / slice header."""
    from make_golden import synth_yuv
    from vvc_intra_b200 import assemble
    enc, cfg = os.path.join(REF, 'EncoderApp'), os.path.join(REF, 'encoder_intra.cfg')
    if not all(os.path.exists(p) for p in (enc, cfg)):
        pytest.skip('oracle/_ref binaries are built only in the container that has /root/reference')
    n = 2
    data = b''
    for f in range(n):
        Y, U, V = synth_yuv(w, h, bits, f)
        data += Y.tobytes() + U.tobytes() + V.tobytes()
    (tmp_path / 'in.yuv').write_bytes(data)
    (tmp_path / 'Time_python.dat').write_bytes(b'')
    args = [enc, '-c', cfg, '-i', 'in.yuv', '-wdt', str(w), '-hgt', str(h), '-q', str(qp), '-fr', '30', '--InputBitDepth=%d' % bits,
            '--InternalBitDepth=%d' % bits, '--OutputBitDepth=%d' % bits] + extra
    procs = [subprocess.Popen(args + ['-f', '1', '--FrameSkip=%d' % f, '-b', 'f%d.bin' % f], cwd=tmp_path, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for f in range(n)]
    procs.append(subprocess.Popen(args + ['-f', str(n), '-b', 'seq.bin'], cwd=tmp_path, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    assert all(p.wait() == 0 for p in procs)
    assemble.assemble_sequential([str(tmp_path / ('f%d.bin' % f)) for f in range(n)], str(tmp_path / 'all.bin'))
    assert (tmp_path / 'all.bin').read_bytes() == (tmp_path / 'seq.bin').read_bytes()
    assemble.gather_sequential([str(tmp_path / ('f%d.bin' % f)) for f in range(n)], str(tmp_path / 'lib.bin'))      # the library's C++ gather
    assert (tmp_path / 'lib.bin').read_bytes() == (tmp_path / 'seq.bin').read_bytes()
