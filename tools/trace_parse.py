"""Parser for the binary trace written by oracle/ref_trace_hooks.cpp (golden-vector generator).

TEST INFRASTRUCTURE: used by tools/make_golden.py (in the build container, where the reference
encoder exists) and by tests/ to read the committed fixtures.  Record layouts are documented next
to each emit() in oracle/ref_trace_hooks.cpp.
"""
import struct
import numpy as np


class _Cur:
    def __init__(self, b):
        self.b, self.o = b, 0

    def i32(self, n=None):
        if n is None:
            v = struct.unpack_from('<i', self.b, self.o)[0]
            self.o += 4
            return v
        v = np.frombuffer(self.b, '<i4', n, self.o).copy()
        self.o += 4 * n
        return v

    def u32(self, n=None):
        if n is None:
            v = struct.unpack_from('<I', self.b, self.o)[0]
            self.o += 4
            return v
        v = np.frombuffer(self.b, '<u4', n, self.o).copy()
        self.o += 4 * n
        return v

    def take(self, nbytes):
        v = bytes(self.b[self.o:self.o + nbytes])
        self.o += nbytes
        return v

    def u64(self):
        v = struct.unpack_from('<Q', self.b, self.o)[0]
        self.o += 8
        return v

    def f64(self):
        v = struct.unpack_from('<d', self.b, self.o)[0]
        self.o += 8
        return v

    def i16(self, n):
        v = np.frombuffer(self.b, '<i2', n, self.o).copy()
        self.o += 2 * n
        return v

    def done(self):
        assert self.o == len(self.b), (self.o, len(self.b))


def _modes(c, with_cost=True, with_isp=False):
    n = c.i32()
    out = []
    for _ in range(n):
        mip, mrl = c.i32(), c.i32()
        isp = c.i32() if with_isp else 0
        mode = c.i32()
        cost = c.f64() if with_cost else 0.0
        out.append(dict(mip=mip, mrl=mrl, isp=isp, mode=mode, cost=cost))
    return out


def parse_record(tag, payload):
    c = _Cur(payload)
    r = dict(tag=tag)
    if tag == 'V':
        r['visit'] = c.u32()
        for k in ('poc', 'x', 'y', 'w', 'h', 'lfnst', 'mts', 'bd', 'qp'):
            r[k] = c.i32()
        r['sqrt_lambda'] = c.f64()
        r['mpm'] = c.i32(6)
        r['num_cand_mpm'] = c.i32()
        r['mip_ctx'] = c.i32()
        # mipFlag[2], mrl0[2], mrl1[2], isp0, mpmFlag[2], planarFlag[2]
        r['rates'] = c.u32(11)
        r['pic_w'], r['pic_h'] = c.i32(), c.i32()
        r['org'] = c.i16(r['w'] * r['h']).reshape(r['h'], r['w'])
    elif tag == 'L':
        r['visit'] = c.u32()
        r['variant'] = c.i32()
        r['rd'] = _modes(c)
        r['had'] = _modes(c)
        r['final'] = _modes(c, with_cost=False, with_isp=True)
    elif tag == 'R':
        r['visit'] = c.u32()
        for k in ('mrl', 'force', 'w', 'h', 'avail_al', 'n_above', 'n_above_right', 'n_left', 'n_below_left', 'has_filt'):
            r[k] = c.i32()
        w, h, mrl = r['w'], r['h'], r['mrl']
        r['unf_top'] = c.i16(2 * w + 1 + mrl)
        r['unf_left'] = c.i16(2 * h + 1 + mrl)
        if r['has_filt']:
            r['filt_top'] = c.i16(2 * w + 1 + mrl)
            r['filt_left'] = c.i16(2 * h + 1 + mrl)
        r['reco_top'] = c.i16(4 * (2 * w + 8)).reshape(4, 2 * w + 8)
        r['reco_left'] = c.i16((2 * h + 4) * 4).reshape(2 * h + 4, 4)
    elif tag == 'P':
        r['visit'] = c.u32()
        for k in ('mip', 'mode', 'mrl', 'w', 'h', 'is_ver', 'ref_filter', 'interp', 'pdpc', 'angle', 'inv_angle', 'ang_scale'):
            r[k] = c.i32()
        r['sad'], r['satd'], r['hash'] = c.u64(), c.u64(), c.u64()
        if c.i32():
            r['pred'] = c.i16(r['w'] * r['h']).reshape(r['h'], r['w'])
    elif tag == 'B':
        r['visit'] = c.u32()
        r['mip'], r['mode'], r['mrl'] = c.i32(), c.i32(), c.i32()
        r['bits'] = c.u64()
    elif tag == 'S':
        for k in ('w', 'h', 'bd', 'max_cand'):
            r[k] = c.i32()
        n = c.i32()
        r['resi'] = c.i16(r['w'] * r['h']).reshape(r['h'], r['w'])
        r['modes'] = []
        for _ in range(n):
            mts, sel = c.i32(), c.i32()
            r['modes'].append(dict(mts=mts, selected=sel, coeff=c.i32(r['w'] * r['h']).reshape(r['h'], r['w'])))
    elif tag == 'Q':
        for k in ('w', 'h', 'bd', 'mts', 'lfnst', 'load_tr', 'qp', 'per', 'rem', 'abs_sum', 'dep_quant'):
            r[k] = c.i32()
        r['lambda'] = c.f64()
        n = r['w'] * r['h']
        r['resi'] = c.i16(n).reshape(r['h'], r['w'])
        r['coeff'] = c.i32(n).reshape(r['h'], r['w'])
        r['level'] = c.i32(n).reshape(r['h'], r['w'])
    elif tag == 'I':
        for k in ('w', 'h', 'bd', 'mts', 'qp', 'per', 'rem'):
            r[k] = c.i32()
        n = r['w'] * r['h']
        r['level'] = c.i32(n).reshape(r['h'], r['w'])
        r['resi'] = c.i16(n).reshape(r['h'], r['w'])
    elif tag == 'D':
        for k in ('w', 'h', 'bd', 'mts', 'lfnst', 'qp', 'per', 'rem', 'abs_sum', 'cbf_delta'):
            r[k] = c.i32()
        r['lambda'] = c.f64()
        r['rates'] = c.u32(2 * (2 + 36 + 63 + 40))
        n = r['w'] * r['h']
        r['resi'] = c.i16(n).reshape(r['h'], r['w'])
        r['coeff'] = c.i32(n).reshape(r['h'], r['w'])
        r['level'] = c.i32(n).reshape(r['h'], r['w'])
    elif tag == 'C':
        for k in ('w', 'h', 'mts', 'ts_allowed', 'mts_allowed', 'dep_quant'):
            r[k] = c.i32()
        r['bits'] = c.u64()
        r['states'] = np.frombuffer(c.take(2 * 3 * 174), '<u2').reshape(174, 3).copy()
        r['level'] = c.i32(r['w'] * r['h']).reshape(r['h'], r['w'])
    elif tag == 'F':
        for k in ('w', 'h', 'bd', 'mts', 'lfnst', 'intra_mode', 'qp', 'per', 'rem', 'abs_sum', 'cbf_delta'):
            r[k] = c.i32()
        r['lambda'] = c.f64()
        r['rates'] = c.u32(2 * (2 + 36 + 63 + 40))
        n = r['w'] * r['h']
        r['resi'] = c.i16(n).reshape(r['h'], r['w'])
        r['coeff'] = c.i32(n).reshape(r['h'], r['w'])
        r['level'] = c.i32(n).reshape(r['h'], r['w'])
    elif tag == 'J':
        for k in ('w', 'h', 'bd', 'mts', 'lfnst', 'intra_mode', 'qp'):
            r[k] = c.i32()
        n = r['w'] * r['h']
        r['level'] = c.i32(n).reshape(r['h'], r['w'])
        r['resi'] = c.i16(n).reshape(r['h'], r['w'])
    elif tag == 'T':
        for k in ('w', 'h', 'bd', 'qp', 'per', 'rem', 'abs_sum'):
            r[k] = c.i32()
        r['lambda'] = c.f64()
        r['rates'] = c.u32(2 * 22)
        n = r['w'] * r['h']
        r['resi'] = c.i16(n).reshape(r['h'], r['w'])
        r['coeff'] = c.i32(n).reshape(r['h'], r['w'])
        r['level'] = c.i32(n).reshape(r['h'], r['w'])
    elif tag == 'H':
        r['w'], r['h'], r['result'] = c.i32(), c.i32(), c.i32()
        r['org'] = c.i16(r['w'] * r['h']).reshape(r['h'], r['w'])
    else:
        raise ValueError('unknown tag %r' % tag)
    c.done()
    return r


def iter_records(path_or_bytes):
    b = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, 'rb').read()
    o = 0
    while o < len(b):
        tag = chr(b[o])
        n = struct.unpack_from('<I', b, o + 1)[0]
        yield parse_record(tag, b[o + 5:o + 5 + n])
        o += 5 + n


def group_visits(records):
    """Group V/R/P/B/L records into per-visit dicts; returns (visits, tu_records)."""
    visits, tus, cur = [], [], None
    for r in records:
        t = r['tag']
        if t == 'V':
            cur = dict(head=r, refs=[], evals=[], lists=None)
            visits.append(cur)
        elif t == 'R':
            cur['refs'].append(r)
        elif t == 'P':
            r['ref_idx'] = len(cur['refs']) - 1
            cur['evals'].append(r)
        elif t == 'B':
            e = cur['evals'][-1]
            assert (e['mip'], e['mode'], e['mrl']) == (r['mip'], r['mode'], r['mrl']), (e['mode'], r['mode'])
            e['bits'] = r['bits']
        elif t == 'L':
            cur['lists'] = r
        else:
            tus.append(r)
    return visits, tus


if __name__ == '__main__':
    import sys
    import collections
    cnt = collections.Counter()
    recs = list(iter_records(sys.argv[1]))
    for r in recs:
        cnt[r['tag']] += 1
    print(dict(cnt))
    visits, tus = group_visits(recs)
    shapes = collections.Counter((v['head']['w'], v['head']['h']) for v in visits)
    print('visits', len(visits), dict(shapes))
    v = visits[0]
    print({k: v['head'][k] for k in v['head'] if k != 'org'})
    print([(e['mip'], e['mrl'], e['mode'], e['sad'], e['satd'], e.get('bits')) for e in v['evals']][:12])
    print(v['lists'])
