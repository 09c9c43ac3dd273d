"""Frame-parallel gather (vvc_intra_b200/assemble.py): pictures encoded by independent encoder processes, concatenated, decoded by the reference decoder.
CPU test on the reference binaries built by oracle/Makefile.ref (skipped where they are absent)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle/_ref')


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
    with pytest.raises(ValueError):
        (tmp_path / 'bad.bin').write_bytes(b'\x00\x00\x01\x11\x02')
        assemble.concat_segments([str(tmp_path / 'bad.bin')], str(tmp_path / 'x.bin'))
