#!/usr/bin/env python
"""One served encoder (no broker) on a 128x128 10-bit crop next to the plain one, then again with per-stage kernel timing; prints the
shim reports.  GPU box only; a measuring aid for profiles/."""
import json
import os
import resource
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tools')]
import bench

qp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
d = tempfile.mkdtemp(prefix='vvcdirect_')
bench.write_crop(d, [bench.synth_luma(k) for k in range(2)], 1)
ref = os.path.join(ROOT, 'oracle/_ref')
cfg = os.path.join(ref, 'encoder_intra.cfg')


def run(exe, out, env=None):
    r0 = resource.getrusage(resource.RUSAGE_CHILDREN)
    t0 = time.perf_counter()
    subprocess.run(bench.encoder_cmd(os.path.join(ref, exe), cfg, qp, out), cwd=d, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
    return dict(wall=time.perf_counter() - t0, user=r1.ru_utime - r0.ru_utime, sys=r1.ru_stime - r0.ru_stime)


print('plain ', run('EncoderApp', 'plain.bin'))
print('served', run('EncoderAppServe', 's.bin', dict(os.environ, VVCB_SHIM_REPORT=os.path.join(d, 'rep.json'))))
print(open(os.path.join(d, 'rep.json')).read())
print('identical', open(os.path.join(d, 'plain.bin'), 'rb').read() == open(os.path.join(d, 's.bin'), 'rb').read())
print('timed ', run('EncoderAppServe', 's2.bin', dict(os.environ, VVCB_SHIM_REPORT=os.path.join(d, 'rep2.json'), VVCB_SHIM_KERNEL_TIMES='1')))
r = json.load(open(os.path.join(d, 'rep2.json')))
print({k: r[k] for k in ('tu_kernel_ms', 'tu_calls', 'rmd_kernel_ms', 'rmd_calls', 'engine_wait_s', 'visits', 'jobs_prefetched')})
