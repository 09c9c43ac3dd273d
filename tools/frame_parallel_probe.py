#!/usr/bin/env python
"""Measuring aid (GPU box): frame-parallel bit-exact encode of a synthetic sequence over the GPUs of the box against the plain sequential encoder.

    python tools/frame_parallel_probe.py --devices 0,1 --frames 8 --size 256x128 --qp 32 > gpurun_out/frame_parallel.json

Runs `oracle/_ref/EncoderApp -f N` (the plain reference, one process) and vvc_intra_b200.frame_parallel.encode_sequence (one served encoder process per
picture, picture k behind the broker of GPU k mod N), compares bitstream and reconstruction byte for byte, prints one JSON line with both wall times and the
brokers' counters.  VVCB_PROBE_BROKER / VVCB_PROBE_LIBDIR point it at the CPU stand-in (tests/fake_engine) for development without a GPU."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from make_golden import synth_yuv  # noqa: E402
from vvc_intra_b200 import frame_parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--devices', default='0')
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--size', default='256x128')
    ap.add_argument('--bits', type=int, default=10)
    ap.add_argument('--qp', type=int, default=32)
    ap.add_argument('--max-procs', type=int, default=None)
    a = ap.parse_args()
    w, h = (int(v) for v in a.size.split('x'))
    ref = os.path.join(ROOT, 'oracle/_ref')
    tmp = tempfile.mkdtemp(prefix='vvcfp_')
    with open(os.path.join(tmp, 'in.yuv'), 'wb') as f:
        for k in range(a.frames):
            for p in synth_yuv(w, h, a.bits, k):
                f.write(p.tobytes())
    open(os.path.join(tmp, 'Time_python.dat'), 'w').close()
    args = ['-c', os.path.join(ref, 'encoder_intra.cfg'), '-i', 'in.yuv', '-wdt', str(w), '-hgt', str(h), '-q', str(a.qp), '-fr', '30',
            '--InputBitDepth=%d' % a.bits, '--InternalBitDepth=%d' % a.bits, '--OutputBitDepth=%d' % a.bits]
    env = dict(os.environ)
    env.pop('VVCB_BROKER', None)
    if os.environ.get('VVCB_PROBE_LIBDIR'):
        env['LD_LIBRARY_PATH'] = os.environ['VVCB_PROBE_LIBDIR']
    t0 = time.perf_counter()
    seq = subprocess.Popen([os.path.join(ref, 'EncoderApp')] + args + ['-f', str(a.frames), '-b', 'seq.bin', '-o', 'seq.yuv'], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    devices = [int(d) for d in a.devices.split(',')]
    st = frame_parallel.encode_sequence(os.path.join(ref, 'EncoderAppServe'), args, a.frames, os.path.join(tmp, 'all.bin'), devices=devices, bit_depth=a.bits, frame_size=(w, h),
                                        workdir=os.path.join(tmp, 'work'), max_procs=a.max_procs, broker_bin=os.environ.get('VVCB_PROBE_BROKER'), env=env,
                                        recon_path=os.path.join(tmp, 'all.yuv'), cwd=tmp)
    if seq.wait() != 0:
        raise SystemExit('the plain encoder failed')
    seq_s = time.perf_counter() - t0                  # both ran side by side on the host cores: an upper bound for the plain encoder alone
    same = [open(os.path.join(tmp, x), 'rb').read() == open(os.path.join(tmp, y), 'rb').read() for x, y in (('all.bin', 'seq.bin'), ('all.yuv', 'seq.yuv'))]
    out = {'frames': a.frames, 'size': a.size, 'bits': a.bits, 'qp': a.qp, 'devices': devices, 'bitstream_identical': same[0], 'reconstruction_identical': same[1],
           'bytes': st['bytes'], 'frame_parallel_seconds': st['seconds'], 'sequential_plain_seconds_alongside': seq_s,
           'pictures': st['pictures'], 'brokers': [{k: d.get(k) for k in ('device', 'pictures', 'clients_seen', 'requests', 'cycles', 'visits', 'tu_jobs', 'kernel_launches', 'max_batch')} for d in st['devices']]}
    print(json.dumps(out))
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    if not all(same):
        raise SystemExit('frame-parallel gather differs from the sequential encoder')


if __name__ == '__main__':
    main()
