"""Frame-parallel assembly: all-intra pictures are independent coded video sequences, so the pictures of a sequence can be
encoded by independent encoder processes (one per picture: `-f 1 --FrameSkip=<n>`, on any GPU of the box through the broker) and gathered afterwards.

Three gathers, from the plainest to the bit-exact one:

* `concat_segments`: each per-picture bitstream carries its own parameter sets and one IDR picture with POC 0; their plain concatenation is a conforming
  bitstream of consecutive coded video sequences that decodes to exactly the pictures the encoders reconstructed.  It is NOT byte-identical to the
  sequential encoder's bitstream: that one codes pictures 1.. as CRA with POC = picture number (NAL unit type, slice_pic_order_cnt_lsb and the
  reference-picture-list bits of the slice header differ).  `diff_against_sequential` reports exactly which NAL units differ.
* `assemble_sequential`: the bit-exact gather.  Re-writes the slice NAL unit of picture n >= 1 into what the sequential encoder writes for it -- CRA,
  slice_pic_order_cnt_lsb = n, the reference-picture-list bits of a non-IDR picture, header re-aligned, emulation prevention redone -- so that the
  gathered stream is byte-identical to `EncoderApp -f <N>` (tests/test_assemble.py).  Nothing else differs: every CRA picture of an all-intra sequence
  has pending-RAS initialisation (EL/EncGOP.cpp:4213-4225), which resets the ALF APS ids (EL/EncAdaptiveLoopFilter.cpp:667-674) and the SAO state
  exactly as an IDR does, and the parameter sets are re-sent with every IRAP picture (EL/EncGOP.cpp:2754-2759, ReWriteParamSets).  The LMCS analysis
  of an intra picture starts from that picture's own statistics with its state re-initialised (EL/EncReshape.cpp:564-789); the header reader steps over
  slice_lmcs_aps_id, but no test stream has slice_lmcs_enabled_flag = 1 -- the reference switches LMCS off for all the synthetic content tried.
* `parcat_segments`: what the reference's own APP/Parcat does with random-access segments that overlap by their IDR picture (parcat.cpp:247-384):
  parameter sets and the IDR picture of segments 2.. dropped, the POC of the other pictures advanced by the pictures gathered so far.  Checked byte for
  byte against the reference's Parcat binary.

The syntax reader is vvc_intra_b200/hls.py."""
import os

from . import hls
from .hls import split_nal_units, nal_unit_type, NAL_NAMES          # re-exported: the gather statistics name NAL unit types


def concat_segments(paths, out_path):
    """The gather: per-picture bitstreams in picture order -> one bitstream.  Returns per-segment statistics (bytes, NAL unit types)."""
    stats = []
    with open(out_path, 'wb') as out:
        for p in paths:
            data = open(p, 'rb').read()
            units = split_nal_units(data)
            if not any(nal_unit_type(u) in (24, 25) for _, u in units):
                raise ValueError('%s holds no IDR picture: not a self-contained coded video sequence' % p)
            out.write(data)
            stats.append({'path': os.path.basename(p), 'bytes': len(data), 'nal_units': [NAL_NAMES.get(nal_unit_type(u), str(nal_unit_type(u))) for _, u in units]})
    return stats


def _gather_call(name, paths, out_path, *extra):
    """The library's gather entry points (include/vvc_intra_b200_gather.h): host code of libvvc_intra_b200.so, no CUDA device involved."""
    import ctypes as C
    from . import engine
    lib = engine.load_library()
    arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    err = C.create_string_buffer(512)
    rc = getattr(lib, name)(arr, len(paths), os.fsencode(out_path), *extra, err, len(err))
    if rc == -3:
        raise NotImplementedError(err.value.decode())
    if rc != 0:
        raise ValueError(err.value.decode())


def gather_sequential(paths, out_path, rewrite_param_sets=True):
    """vvcb_gather_sequential: the bit-exact gather done by the library (C++; `assemble_sequential` below is its Python twin, kept as the cross-check
    of the tests and for the per-picture statistics).  Returns the number of bytes written."""
    import ctypes as C
    n = C.c_uint64(0)
    _gather_call('vvcb_gather_sequential', list(paths), out_path, C.c_int(1 if rewrite_param_sets else 0), C.byref(n))
    return n.value


def gather_parcat(paths, out_path):
    """vvcb_gather_parcat: the reference's Parcat done by the library (twin: `parcat_segments`).  Returns the number of pictures re-numbered."""
    import ctypes as C
    k = C.c_int(0)
    _gather_call('vvcb_gather_parcat', list(paths), out_path, C.byref(k))
    return k.value


def diff_against_sequential(assembled, sequential):
    """NAL-by-NAL comparison of an assembled bitstream with the sequential encoder's: list of (index, type assembled, type sequential, bytes assembled,
    bytes sequential, number of differing bytes) for the units that differ."""
    a, b = split_nal_units(assembled), split_nal_units(sequential)
    if len(a) != len(b):
        return [('count', len(a), len(b))]
    out = []
    for i, ((_, ua), (_, ub)) in enumerate(zip(a, b)):
        if ua != ub:
            n = sum(x != y for x, y in zip(ua, ub)) + abs(len(ua) - len(ub))
            out.append((i, NAL_NAMES.get(nal_unit_type(ua), str(nal_unit_type(ua))), NAL_NAMES.get(nal_unit_type(ub), str(nal_unit_type(ub))), len(ua), len(ub), n))
    return out


class _ParameterSets:
    """The parameter sets seen so far in a stream (the slice header's syntax depends on them)."""
    def __init__(self):
        self.sps, self.pps = {}, {}

    def see(self, unit):
        t = nal_unit_type(unit)
        if t == hls.NAL_SPS:
            s = hls.parse_sps(hls.unescape(unit[2:]))
            self.sps[s['sps_id']] = s
        elif t == hls.NAL_PPS:
            p = hls.parse_pps(hls.unescape(unit[2:]), self.sps)
            self.pps[p['pps_id']] = p

    def slice_header(self, unit):
        rbsp = hls.unescape(unit[2:])
        r = hls.BitReader(rbsp)
        if nal_unit_type(unit) in (hls.NAL_IDR_W_RADL, hls.NAL_IDR_N_LP, hls.NAL_CRA):
            r.flag()
        pps = self.pps[r.ue()]
        sps = self.sps[pps['sps_id']]
        return rbsp, sps, hls.parse_intra_slice_header(rbsp, nal_unit_type(unit), sps, self.pps)


def renumber_idr_as_cra(unit, poc, sets):
    """The slice NAL unit of an IDR picture (POC 0) -> the NAL unit the sequential all-intra encoder writes for the same picture at position `poc` >= 1:
    NAL_UNIT_CODED_SLICE_CRA, slice_pic_order_cnt_lsb = poc, then the syntax a non-IDR picture carries between the POC and the SAO flags
    (EL/VLCWriter.cpp:1238-1335): ref_pic_list_sps_flag[0] = 1 with ref_pic_list_idx[0] = 0 -- the list of GOP position 0, EL/EncLib.cpp:1632 -- the same
    for list 1 when the PPS signals it separately, and slice_temporal_mvp_enabled_flag = 1 when the SPS enables it (EL/EncGOP.cpp:2226).  The rest of
    the header moves by those bits, is re-aligned (EL/EncGOP.cpp:4282), and the slice data follows unchanged."""
    rbsp, sps, h = sets.slice_header(unit)
    if nal_unit_type(unit) not in (hls.NAL_IDR_W_RADL, hls.NAL_IDR_N_LP) or h['poc_lsb'] != 0:
        raise ValueError('not the IDR picture of a one-picture segment')
    if sps['idr_rpl_present']:
        raise NotImplementedError('sps_idr_rpl_present_flag: the IDR header already carries reference picture lists')
    if sps['num_rpl0'] < 1 or sps['num_rpl1'] < 1:
        raise NotImplementedError('no reference picture list structure in the SPS')
    w = hls.BitWriter()
    w.copy(rbsp, 0, h['poc_pos'])
    w.u(poc & ((1 << sps['poc_bits']) - 1), sps['poc_bits'])
    w.u(1, 1)
    if sps['num_rpl0'] > 1:
        w.u(0, (sps['num_rpl0'] - 1).bit_length())
    pps = sets.pps[hls.BitReader(rbsp, 1).ue()]
    if pps['rpl1_idx_present']:
        w.u(1, 1)
        if sps['num_rpl1'] > 1:
            w.u(0, (sps['num_rpl1'] - 1).bit_length())
    if sps['temporal_mvp']:
        w.u(1, 1)
    w.copy(rbsp, h['after_rpl'], h['header_end'])
    w.align()
    return hls.with_nal_unit_type(unit[:2], hls.NAL_CRA) + hls.escape(w.tobytes() + rbsp[h['data_start']:])


def assemble_sequential(paths, out_path, rewrite_param_sets=True):
    """The bit-exact gather: one-picture bitstreams in picture order -> the bitstream the sequential encoder writes for the whole sequence.
    `rewrite_param_sets` follows the encoder's ReWriteParamSets (1 in the shipped configuration: every IRAP picture re-sends VPS / SPS / PPS);
    with 0 the parameter sets of pictures 1.. are dropped.  Returns per-picture statistics."""
    stats = []
    with open(out_path, 'wb') as out:
        for n, p in enumerate(paths):
            data = open(p, 'rb').read()
            units = split_nal_units(data)
            vcl = [nal_unit_type(u) for _, u in units if 8 <= nal_unit_type(u) < 16 or 24 <= nal_unit_type(u) < 28]
            if len(vcl) != 1 or vcl[0] not in (hls.NAL_IDR_W_RADL, hls.NAL_IDR_N_LP):
                raise ValueError('%s is not a one-picture segment (one IDR slice)' % p)
            sets, end, written = _ParameterSets(), 0, 0
            for off, u in units:
                t = nal_unit_type(u)
                sets.see(u)
                prefix, end = data[end:off], off + len(u)
                if n > 0 and t in (hls.NAL_IDR_W_RADL, hls.NAL_IDR_N_LP):
                    u = renumber_idr_as_cra(u, n, sets)
                elif n > 0 and not rewrite_param_sets and t in (hls.NAL_DPS, hls.NAL_VPS, hls.NAL_SPS, hls.NAL_PPS):
                    continue
                out.write(prefix + u)
                written += len(prefix) + len(u)
            stats.append({'path': os.path.basename(p), 'poc': n, 'bytes_in': len(data), 'bytes_out': written})
    return stats


def parcat_segments(paths, out_path, bits_for_poc=8):
    """The reference's APP/Parcat (parcat.cpp:247-384, main :418-446) on the same inputs, byte for byte: segments in order; of segments 2.. the
    parameter sets / access unit delimiters ahead of the IDR picture, the IDR picture itself and the suffix SEI that follows it are dropped (the
    segments overlap by that picture); every other slice has `poc_base` -- the non-IDR pictures gathered so far -- added to its
    slice_pic_order_cnt_lsb.  Like the reference tool it takes the POC field to be `bits_for_poc` = 8 bits wide (parcat.cpp:263) and keeps the old
    low bit when it merges the new value (:337-339 masks one bit too many) -- harmless whenever poc_base is even or the sum has that bit set, and
    reproduced here because the contract is the tool's output.  Returns the number of pictures re-numbered."""
    sets, poc_base, mask = _ParameterSets(), 0, (1 << bits_for_poc) - 1
    with open(out_path, 'wb') as out:
        for idx, p in enumerate(paths, start=1):
            data = open(p, 'rb').read()
            end, cnt, idr_found, skip_next_sei = 0, 0, False, False
            for off, u in split_nal_units(data):
                t = nal_unit_type(u)
                sets.see(u)
                prefix, end = data[end:off], off + len(u)
                is_idr = t in (hls.NAL_IDR_W_RADL, hls.NAL_IDR_N_LP)
                if 7 < t < 15 or t == hls.NAL_CRA:
                    _, sps, h = sets.slice_header(u)
                    if sps['poc_bits'] != bits_for_poc:
                        raise NotImplementedError('the reference tool only handles an 8-bit slice_pic_order_cnt_lsb')
                    pos = 16 + h['poc_pos']                      # the tool counts bits from the NAL unit header on and edits the escaped bytes in place
                    byte, hi = pos >> 3, pos & 7
                    word = (u[byte] << 8) | u[byte + 1]
                    low_bits = 16 - hi - bits_for_poc
                    new_lsb = (((word >> low_bits) & 0xff) + poc_base + (1 << bits_for_poc)) & mask
                    word = ((word >> (16 - hi)) << (16 - hi)) | (new_lsb << low_bits) | (word & ((1 << (low_bits + 1)) - 1))
                    u = u[:byte] + bytes([word >> 8, word & 0xff]) + u[byte + 2:]
                    cnt += 1                                     # one slice per picture here: every slice is the first of its picture
                if idx > 1 and is_idr:
                    skip_next_sei = idr_found = True
                drop = (idx > 1 and is_idr) or (idx > 1 and not idr_found and t in (hls.NAL_DPS, hls.NAL_VPS, hls.NAL_SPS, hls.NAL_PPS, hls.NAL_APS, hls.NAL_AUD)) \
                    or (t == hls.NAL_SUFFIX_SEI and skip_next_sei)
                if not drop:
                    out.write(prefix + u)
                if t == hls.NAL_SUFFIX_SEI and skip_next_sei:
                    skip_next_sei = False
            poc_base += cnt
    return poc_base


def main(argv=None):
    """python -m vvc_intra_b200.assemble [--parcat | --concat] [--no-param-sets] seg0.bin seg1.bin ... out.bin
    default: the bit-exact gather of one-picture segments; --parcat: the reference tool's behaviour (same argument order as `Parcat`)."""
    import argparse
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument('--parcat', action='store_true', help='overlapping random-access segments, as APP/Parcat')
    ap.add_argument('--concat', action='store_true', help='plain concatenation of self-contained segments')
    ap.add_argument('--no-param-sets', action='store_true', help='bit-exact gather for ReWriteParamSets=0: parameter sets only ahead of the first picture')
    ap.add_argument('files', nargs='+', help='segments in order, then the output file')
    a = ap.parse_args(argv)
    if len(a.files) < 2:
        ap.error('need at least one segment and the output file')
    segs, out = a.files[:-1], a.files[-1]
    if a.parcat:
        print('%d pictures re-numbered' % gather_parcat(segs, out))
    elif a.concat:
        print('%d segments, %d bytes' % (len(segs), sum(s['bytes'] for s in concat_segments(segs, out))))
    else:
        print('%d pictures, %d bytes' % (len(segs), gather_sequential(segs, out, rewrite_param_sets=not a.no_param_sets)))
    return 0


if __name__ == '__main__':
    import sys
    sys.exit(main())
