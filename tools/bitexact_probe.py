#!/usr/bin/env python
"""Runs bench.py's bit-exact leg alone for a few (walkers per core, broker workers) settings; GPU box only, a measuring aid for profiles/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tools')]
import bench

frames = [bench.synth_luma(f) for f in range(bench.NFRAMES)]
out = []
for over, workers in [tuple(int(x) for x in a.split(',')) for a in sys.argv[1:]]:
    os.environ['VVCB_BENCH_OVERSUBSCRIBE'], os.environ['VVCB_BENCH_WORKERS'] = str(over), str(workers)
    r = bench.bitexact_leg(frames, 0)
    r.pop('note', None), r.pop('workload', None)
    out.append(dict(over=over, workers=workers, **r))
    print(json.dumps(out[-1]), flush=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'bitexact_probe.json'), 'w'), indent=1)
