"""Frame-parallel encode of an all-intra sequence (north_star's multi-GPU layout): one encoder process per picture, the pictures
dealt round-robin over the GPUs of the box (shard.shard_units), the encoder processes of one GPU sharing its engine context through a broker
(include/vvc_intra_b200_broker.h), and the per-picture bitstreams gathered into the bitstream the sequential encoder writes for the whole sequence
(assemble.assemble_sequential: byte-identical to `EncoderApp -f N`).  Nothing is exchanged between the GPUs: all-intra pictures are independent
(IntraPeriod 1, GOPSize 1: BIN/encoder_intra.cfg:26-28), which is also why the reference's own tooling concatenates separately encoded segments
(APP/Parcat).

The encoder is the reference's encoder linked against libvvc_intra_b200.so as INTEGRATION.md section 8 shows (in this repository's tests:
oracle/_ref/EncoderAppServe); this module only starts processes and gathers files -- it holds no arithmetic of the path.

    python -m vvc_intra_b200.frame_parallel --encoder ./EncoderAppServe --frames 64 --devices 0,1,2,3 --bit-depth 10 --size 3840x2160 \\
        --out all.bin -- -c encoder_intra.cfg -i in.yuv -wdt 3840 -hgt 2160 -q 32 -fr 30 --InputBitDepth=10 --InternalBitDepth=10
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

from . import assemble, shard

PKG = os.path.dirname(os.path.abspath(__file__))


class FrameParallelError(RuntimeError):
    pass


def _strip_option(args, names):
    """Encoder arguments without the options this driver sets itself (`-f N`, `--FramesToBeEncoded=N`, ...)."""
    out, skip = [], False
    for a in args:
        if skip:
            skip = False
            continue
        if a in names:
            skip = True
            continue
        if any(a.startswith(n + '=') for n in names if n.startswith('--')):
            continue
        out.append(a)
    return out


def encode_sequence(encoder, encoder_args, n_frames, out_path, devices=(0,), bit_depth=10, frame_size=(1920, 1080), workdir=None, max_procs=None,
                    broker_bin=None, broker_workers=2, env=None, recon_path=None, first_frame=0, rewrite_param_sets=True, timeout=None, cwd=None):
    """Encodes pictures first_frame .. first_frame + n_frames - 1 of the input named in `encoder_args` with one `encoder` process per picture and
    gathers them into `out_path`.  `devices`: the GPUs to use, one broker (one engine context) each; picture k goes to devices[k % len(devices)].
    `max_procs` bounds the encoder processes alive at a time (default: the host's cores).  `recon_path`: also gather the reconstructions (needs no
    `-o` in `encoder_args`).  `cwd`: working directory of the encoder processes.  Returns a dict of statistics (per picture: device, bytes, seconds; per device: the broker's counters)."""
    if n_frames < 1 or not devices:
        raise ValueError('nothing to encode')
    broker_bin = broker_bin or os.path.join(PKG, 'vvcb_broker')
    if not os.path.exists(broker_bin):
        raise FrameParallelError('%s is missing: build the package first (__graft_entry__.build)' % broker_bin)
    args = _strip_option(list(encoder_args), ('-f', '--FramesToBeEncoded', '-fs', '--FrameSkip', '-b', '--BitstreamFile') + (('-o', '--ReconFile') if recon_path else ()))
    env = dict(os.environ if env is None else env)
    max_procs = max_procs or os.cpu_count() or 1
    own_dir = workdir is None
    workdir = tempfile.mkdtemp(prefix='vvcb_fp_') if own_dir else workdir
    os.makedirs(workdir, exist_ok=True)
    cwd = cwd or os.getcwd()                                           # where the encoder's relative paths (cfg, input) resolve
    units = list(range(n_frames))
    per_device = [shard.shard_units(units, r, len(devices)) for r in range(len(devices))]
    clients = max(2, min(max_procs, max(len(u) for u in per_device)) + 1)
    brokers, servers, stats = [], [], {'pictures': [None] * n_frames, 'devices': []}
    t0 = time.perf_counter()
    try:
        for r, dev in enumerate(devices):
            path = os.path.join(workdir, 'broker%d.shm' % r)
            servers.append(subprocess.Popen([broker_bin, path, '--device', str(dev), '--bit-depth', str(bit_depth), '--clients', str(clients),
                                             '--frame', '%dx%d' % tuple(frame_size), '--workers', str(broker_workers)], env=env, stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True))
            brokers.append(path)
        owner = {k: r for r, ks in enumerate(per_device) for k in ks}
        running, pending = {}, list(units)

        def reap(block):
            while running:
                for k, (p, started) in list(running.items()):
                    if p.poll() is None:
                        continue
                    out, _ = p.communicate()
                    del running[k]
                    if p.returncode != 0:
                        raise FrameParallelError('encoder of picture %d failed (%d): %s' % (first_frame + k, p.returncode, (out or '')[-2000:]))
                    stats['pictures'][k] = {'picture': first_frame + k, 'device': devices[owner[k]], 'seconds': time.perf_counter() - started,
                                            'bytes': os.path.getsize(os.path.join(workdir, 'pic%06d.bin' % k))}
                    return
                if not block:
                    return
                for s in servers:
                    if s.poll() is not None:
                        raise FrameParallelError('broker exited: %s' % (s.communicate()[0] or '')[-2000:])
                if timeout is not None and time.perf_counter() - t0 > timeout:
                    raise FrameParallelError('timed out after %.0f s' % timeout)
                time.sleep(0.005)

        while pending or running:
            while pending and len(running) < max_procs:
                k = pending.pop(0)
                cmd = [encoder] + args + ['-f', '1', '--FrameSkip=%d' % (first_frame + k), '-b', os.path.join(workdir, 'pic%06d.bin' % k)]
                if recon_path:
                    cmd += ['-o', os.path.join(workdir, 'pic%06d.yuv' % k)]
                running[k] = (subprocess.Popen(cmd, cwd=cwd, env=dict(env, VVCB_BROKER=brokers[owner[k]]), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True),
                              time.perf_counter())
            reap(block=True)
        stats['encode_seconds'] = time.perf_counter() - t0
        for r, path in enumerate(brokers):
            try:
                stats['devices'].append(dict(json.loads(subprocess.check_output([broker_bin, path, '--stats'], env=env)), device=devices[r], pictures=len(per_device[r])))
            except (subprocess.CalledProcessError, ValueError):
                stats['devices'].append({'device': devices[r], 'pictures': len(per_device[r])})
        stats['gather_bytes'] = assemble.gather_sequential([os.path.join(workdir, 'pic%06d.bin' % k) for k in units], out_path, rewrite_param_sets=rewrite_param_sets)
        if recon_path:
            with open(recon_path, 'wb') as out:
                for k in units:
                    out.write(open(os.path.join(workdir, 'pic%06d.yuv' % k), 'rb').read())
        stats['bytes'] = os.path.getsize(out_path)
        stats['seconds'] = time.perf_counter() - t0
        return stats
    finally:
        for k, (p, _) in running.items():
            p.kill()
        for path, s in zip(brokers, servers):
            if s.poll() is None:
                subprocess.run([broker_bin, path, '--stop'], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for s in servers:
            try:
                s.communicate(timeout=30)
            except subprocess.TimeoutExpired:
                s.kill()
        if own_dir:
            import shutil
            shutil.rmtree(workdir, ignore_errors=True)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n\n')[0], formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--encoder', required=True, help='the reference encoder linked against libvvc_intra_b200.so (INTEGRATION.md section 8)')
    ap.add_argument('--frames', type=int, required=True)
    ap.add_argument('--first-frame', type=int, default=0)
    ap.add_argument('--devices', default='0', help='comma-separated GPU indices, one broker each')
    ap.add_argument('--bit-depth', type=int, default=10)
    ap.add_argument('--size', default='1920x1080', help='luma WxH (sizes the brokers\' frame slots)')
    ap.add_argument('--max-procs', type=int, default=None)
    ap.add_argument('--out', required=True)
    ap.add_argument('--recon', default=None)
    ap.add_argument('--workdir', default=None)
    ap.add_argument('encoder_args', nargs=argparse.REMAINDER, help='after --: the arguments of the encoder (cfg, input, size, QP ...)')
    a = ap.parse_args(argv)
    rest = a.encoder_args[1:] if a.encoder_args[:1] == ['--'] else a.encoder_args
    w, h = (int(v) for v in a.size.lower().split('x'))
    st = encode_sequence(a.encoder, rest, a.frames, a.out, devices=[int(d) for d in a.devices.split(',')], bit_depth=a.bit_depth, frame_size=(w, h),
                         workdir=a.workdir, max_procs=a.max_procs, recon_path=a.recon, first_frame=a.first_frame)
    json.dump({k: st[k] for k in ('bytes', 'seconds', 'encode_seconds', 'devices')}, sys.stdout)
    sys.stdout.write('\n')
    return 0


if __name__ == '__main__':
    sys.exit(main())
