"""Frame-parallel assembly: all-intra pictures are independent coded video sequences, so the pictures of a sequence can be
encoded by independent encoder processes (one per picture: `-f 1 --FrameSkip=<n>`, on any GPU of the box through the broker) and gathered afterwards.

Each per-picture bitstream carries its own parameter sets and one IDR picture with POC 0; their plain concatenation is a conforming bitstream of
consecutive coded video sequences that decodes to exactly the pictures the encoders reconstructed (tests/test_assemble.py checks it with the reference
decoder, and that the reconstruction equals the sequential encoder's).  It is NOT byte-identical to the sequential encoder's bitstream: that one codes
pictures 1.. as CRA with POC = picture number (NAL unit type, slice_pic_order_cnt_lsb and the reference-picture-list bits of the slice header differ;
with ALF on, the APS ids advance per picture, EL/EncAdaptiveLoopFilter.cpp:667-674).  `diff_against_sequential` reports exactly which NAL units differ --
rewriting those slice headers is what the reference's APP/Parcat does for its own use case and is the remaining step to a bit-exact gather."""
import os


def split_nal_units(data):
    """Annex-B byte stream -> list of (offset of the NAL header, NAL unit bytes without start code and trailing zero_bytes)."""
    pos, i = [], 0
    while True:
        j = data.find(b'\x00\x00\x01', i)
        if j < 0:
            break
        pos.append(j + 3)
        i = j + 3
    out = []
    for k, p in enumerate(pos):
        e = pos[k + 1] - 3 if k + 1 < len(pos) else len(data)
        unit = data[p:e]
        while k + 1 < len(pos) and unit.endswith(b'\x00'):       # the zero_byte of the next 4-byte start code / trailing_zero_8bits
            unit = unit[:-1]
        out.append((p, unit))
    return out


def nal_unit_type(unit):
    """VTM 6.1 NAL header (CL/NAL.h, DL/NALread.cpp): zero_tid_required_flag(1) nuh_temporal_id_plus1(3) nal_unit_type_lsb(4) | layer id ..."""
    return ((unit[0] >> 7) << 4) | (unit[0] & 0x0f)


NAL_NAMES = {0: 'PPS', 1: 'AUD', 2: 'PREFIX_SEI', 3: 'SUFFIX_SEI', 4: 'APS', 8: 'TRAIL', 16: 'DPS', 17: 'SPS', 18: 'EOS', 19: 'EOB', 20: 'VPS',
             24: 'IDR_W_RADL', 25: 'IDR_N_LP', 26: 'CRA', 27: 'GRA'}


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
