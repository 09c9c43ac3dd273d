#!/usr/bin/env python
"""Generates tests/golden/assemble/: bitstreams of the UNMODIFIED reference encoder (oracle/_ref/EncoderApp) and the output of the reference's own
segment concatenator (oracle/_ref/Parcat, built by oracle/Makefile.ref from APP/Parcat/parcat.cpp) on synthetic input, for the gathers of
vvc_intra_b200/assemble.py.  Runs only in the build container (/root/reference present).

    python tools/make_assemble_golden.py

  pic_256x128_10b_qp27_f<n>.bin   one encoder process per picture (-f 1 --FrameSkip=n), n = 0..3; picture 3 carries an ALF APS
  pic_256x128_10b_qp27_seq.bin    the sequential encoder over the same four pictures (-f 4): what assemble_sequential must reproduce
  pic_416x240_8b_qp32_f<n>.bin    configs[0]'s picture size and QP, n = 0..1: BOTH pictures carry an ALF APS, with id 7 in the sequential stream too
  pic_416x240_8b_qp32_seq.bin     (the APS id counter restarts with every CRA picture: pending-RAS initialisation)
  seg_64x64_8b_qp32_s<k>.bin      three-picture segments overlapping by one picture (-f 3 --FrameSkip=2k), k = 0..2
  seg_64x64_8b_qp32_parcat.bin    Parcat s0 s1 s2: what parcat_segments must reproduce (it also equals the sequential encoder's stream)"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from make_golden import synth_yuv  # noqa: E402

REF = os.path.join(ROOT, 'oracle/_ref')
OUT = os.path.join(ROOT, 'tests/golden/assemble')


def encoder_args(w, h, bits, qp):
    return [os.path.join(REF, 'EncoderApp'), '-c', os.path.join(REF, 'encoder_intra.cfg'), '-i', 'in.yuv', '-wdt', str(w), '-hgt', str(h), '-q', str(qp),
            '-fr', '30', '--InputBitDepth=%d' % bits, '--InternalBitDepth=%d' % bits, '--OutputBitDepth=%d' % bits]


def write_input(tmp, w, h, bits, n):
    data = b''
    for f in range(n):
        Y, U, V = synth_yuv(w, h, bits, f)
        data += Y.tobytes() + U.tobytes() + V.tobytes()
    open(os.path.join(tmp, 'in.yuv'), 'wb').write(data)
    open(os.path.join(tmp, 'Time_python.dat'), 'wb').write(b'')


def run_all(cmds, tmp):
    procs = [subprocess.Popen(c, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
    if any(p.wait() for p in procs):
        raise RuntimeError('encoder failed')


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix='vvcasm_')
    try:
        w, h, bits, qp, n = 256, 128, 10, 27, 4
        write_input(tmp, w, h, bits, n)
        a = encoder_args(w, h, bits, qp)
        run_all([a + ['-f', '1', '--FrameSkip=%d' % f, '-b', 'f%d.bin' % f] for f in range(n)] + [a + ['-f', str(n), '-b', 'seq.bin']], tmp)
        for f in range(n):
            shutil.copy(os.path.join(tmp, 'f%d.bin' % f), os.path.join(OUT, 'pic_256x128_10b_qp27_f%d.bin' % f))
        shutil.copy(os.path.join(tmp, 'seq.bin'), os.path.join(OUT, 'pic_256x128_10b_qp27_seq.bin'))
        w, h, bits, qp, n = 416, 240, 8, 32, 2
        write_input(tmp, w, h, bits, n)
        a = encoder_args(w, h, bits, qp)
        run_all([a + ['-f', '1', '--FrameSkip=%d' % f, '-b', 'g%d.bin' % f] for f in range(n)] + [a + ['-f', str(n), '-b', 'gseq.bin']], tmp)
        for f in range(n):
            shutil.copy(os.path.join(tmp, 'g%d.bin' % f), os.path.join(OUT, 'pic_416x240_8b_qp32_f%d.bin' % f))
        shutil.copy(os.path.join(tmp, 'gseq.bin'), os.path.join(OUT, 'pic_416x240_8b_qp32_seq.bin'))
        w, h, bits, qp, n = 64, 64, 8, 32, 7
        write_input(tmp, w, h, bits, n)
        a = encoder_args(w, h, bits, qp)
        run_all([a + ['-f', '3', '--FrameSkip=%d' % (2 * k), '-b', 's%d.bin' % k] for k in range(3)], tmp)
        subprocess.check_call([os.path.join(REF, 'Parcat'), 's0.bin', 's1.bin', 's2.bin', 'parcat.bin'], cwd=tmp, stdout=subprocess.DEVNULL)
        for k in range(3):
            shutil.copy(os.path.join(tmp, 's%d.bin' % k), os.path.join(OUT, 'seg_64x64_8b_qp32_s%d.bin' % k))
        shutil.copy(os.path.join(tmp, 'parcat.bin'), os.path.join(OUT, 'seg_64x64_8b_qp32_parcat.bin'))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for f in sorted(os.listdir(OUT)):
        print('%8d  %s' % (os.path.getsize(os.path.join(OUT, f)), f))


if __name__ == '__main__':
    main()
