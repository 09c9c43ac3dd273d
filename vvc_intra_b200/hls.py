"""High-level syntax of the reference's bitstreams (VTM 6.1 draft syntax), as far as the frame-parallel gather needs it: NAL units,
emulation prevention, the sequence / picture parameter sets, and the slice header of an intra slice -- enough to find slice_pic_order_cnt_lsb and the
end of the slice header, which is what re-numbering the pictures of independently encoded segments takes.

Follows the order the reference writes the syntax in: EL/NALwrite.cpp:47-140 (NAL header, emulation prevention), EL/VLCWriter.cpp:736-1130 (codeSPS),
:1748-1776 and :1701-1746 (profile_tier_level, constraint flags), :161-206 (ref_pic_list_struct), :208-491 (codePPS), :1167-1699 (codeSliceHeader),
EL/EncGOP.cpp:4279-4290 (byte alignment between header and slice data).  The reader refuses (NotImplementedError) what the all-intra configuration never
writes -- scaling lists, HRD / VUI, several tiles, inter slices -- rather than guessing at it; each parameter-set parse checks that it ends on the
rbsp_trailing_bits, which is what keeps a mis-read flag from going unnoticed."""

NAL_PPS, NAL_AUD, NAL_PREFIX_SEI, NAL_SUFFIX_SEI, NAL_APS = 0, 1, 2, 3, 4
NAL_TRAIL, NAL_DPS, NAL_SPS, NAL_VPS = 8, 16, 17, 20
NAL_IDR_W_RADL, NAL_IDR_N_LP, NAL_CRA = 24, 25, 26
NAL_NAMES = {0: 'PPS', 1: 'AUD', 2: 'PREFIX_SEI', 3: 'SUFFIX_SEI', 4: 'APS', 8: 'TRAIL', 16: 'DPS', 17: 'SPS', 18: 'EOS', 19: 'EOB', 20: 'VPS',
             24: 'IDR_W_RADL', 25: 'IDR_N_LP', 26: 'CRA', 27: 'GRA'}
I_SLICE = 2


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
    """EL/NALwrite.cpp:47-66: zero_tid_required_flag(1) nuh_temporal_id_plus1(3) nal_unit_type_lsb(4) | nuh_layer_id(7) reserved(1)."""
    return ((unit[0] >> 7) << 4) | (unit[0] & 0x0f)


def with_nal_unit_type(unit, t):
    """The same NAL unit under another type (the types the gather exchanges are all >= 16: zero_tid_required_flag stays 1)."""
    return bytes([((t >> 4) << 7) | (unit[0] & 0x70) | (t & 0x0f)]) + unit[1:]


def unescape(payload):
    """NAL payload -> RBSP: drops every emulation_prevention_three_byte (00 00 03)."""
    out, zeros = bytearray(), 0
    for v in payload:
        if zeros >= 2 and v == 3:
            zeros = 0
            continue
        out.append(v)
        zeros = zeros + 1 if v == 0 else 0
    return bytes(out)


def escape(rbsp):
    """RBSP -> NAL payload exactly as EL/NALwrite.cpp:97-137 writes it (including the final 03 after a trailing zero byte)."""
    out, zeros = bytearray(), 0
    for v in rbsp:
        if zeros == 2 and v <= 3:
            out.append(3)
            zeros = 0
        zeros = zeros + 1 if v == 0 else 0
        out.append(v)
    if zeros > 0:
        out.append(3)
    return bytes(out)


class BitReader:
    def __init__(self, rbsp, pos=0):
        self.d, self.pos = rbsp, pos

    def u(self, n):
        v = 0
        for _ in range(n):
            if self.pos >= 8 * len(self.d):
                raise ValueError('read past the end of the RBSP')
            v = (v << 1) | ((self.d[self.pos >> 3] >> (7 - (self.pos & 7))) & 1)
            self.pos += 1
        return v

    def flag(self):
        return self.u(1)

    def ue(self):
        z = 0
        while self.u(1) == 0:
            z += 1
            if z > 32:
                raise ValueError('bad Exp-Golomb code')
        return (1 << z) - 1 + (self.u(z) if z else 0)

    def se(self):
        k = self.ue()
        return (k + 1) >> 1 if k & 1 else -(k >> 1)

    def byte_aligned(self):
        return (self.pos & 7) == 0

    def at_trailing_bits(self):
        """True when what is left is exactly rbsp_trailing_bits: a one, then zeros to the end."""
        n = 8 * len(self.d)
        if self.pos >= n or n - self.pos > 8:
            return False
        save = self.pos
        ok = self.u(1) == 1 and all(self.u(1) == 0 for _ in range(n - self.pos))
        self.pos = save
        return ok


class BitWriter:
    def __init__(self):
        self.bits = []

    def u(self, v, n):
        self.bits += [(v >> (n - 1 - i)) & 1 for i in range(n)]

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(0, n - 1)
        self.u(v, n)

    def copy(self, rbsp, start, end):
        self.bits += [(rbsp[p >> 3] >> (7 - (p & 7))) & 1 for p in range(start, end)]

    def align(self):
        """EL/EncGOP.cpp:4282 writeByteAlignment: a one, then zeros to the byte boundary."""
        self.bits.append(1)
        while len(self.bits) & 7:
            self.bits.append(0)

    def tobytes(self):
        assert len(self.bits) % 8 == 0
        out = bytearray()
        for i in range(0, len(self.bits), 8):
            b = 0
            for x in self.bits[i:i + 8]:
                b = (b << 1) | x
            out.append(b)
        return bytes(out)


def _ref_pic_list_struct(r, long_term, poc_bits, forbid_zero_delta):
    """EL/VLCWriter.cpp:161-206 for lists of short-term pictures.  Returns the number of entries (the gather only needs to step over the structure)."""
    if long_term:
        raise NotImplementedError('long-term reference pictures')
    n = r.ue()
    for _ in range(n):
        if r.ue() + (1 if forbid_zero_delta else 0) > 0:          # abs_delta_poc_st (minus 1 when a zero delta cannot occur: no weighted prediction)
            r.flag()                                               # strp_entry_sign_flag
    return n


def parse_sps(rbsp):
    """EL/VLCWriter.cpp:736-1130 -> dict of the fields the slice header depends on."""
    r, s = BitReader(rbsp), {}
    r.u(4)                                                           # sps_decoding_parameter_set_id
    sub_layers = r.u(3) + 1
    r.u(5)
    r.u(7); r.flag(); r.u(24)                                        # profile_tier_level: profile, tier, sub-profile
    r.u(5); r.u(4); r.u(2)                                           # constraint info: 5 source / intra-only flags, bit depth idc, chroma idc
    r.u(26)                                                          # the 26 tool constraint flags, :1710-1745 (joint Cb-Cr and BDPCM present, PCM removed)
    r.u(8)                                                           # general_level_idc
    present = [r.flag() for _ in range(sub_layers - 1)]
    while not r.byte_aligned():
        r.flag()
    for p in present:
        if p:
            r.u(8)
    s['sps_id'] = r.ue()
    s['chroma_format_idc'] = r.ue()
    if s['chroma_format_idc'] == 3:
        r.flag()
    s['width'], s['height'] = r.ue(), r.ue()
    r.ue(); r.ue(); r.ue()                                           # bit depths, min_qp_prime_ts_minus4
    s['poc_bits'] = r.ue() + 4
    if s['poc_bits'] > 16:
        raise ValueError('log2_max_pic_order_cnt_lsb out of range')
    s['idr_rpl_present'] = r.flag()
    ordering = r.flag()
    for i in range(sub_layers):
        r.ue(); r.ue(); r.ue()
        if not ordering:
            break
    s['long_term_refs'] = r.flag()
    s['rpl1_copy'] = r.flag()
    # the weighted-prediction flags follow the lists in the SPS although the lists' coding depends on them: the reference writes the lists with what
    # the encoder holds and reads them back with the SPS defaults; all-intra streams have both off, and the parse below is checked against that
    start = r.pos
    for forbid in (True, False):
        r.pos = start
        s['num_rpl0'] = r.ue()
        for _ in range(s['num_rpl0']):
            _ref_pic_list_struct(r, s['long_term_refs'], s['poc_bits'], forbid)
        s['num_rpl1'] = s['num_rpl0']
        if not s['rpl1_copy']:
            s['num_rpl1'] = r.ue()
            for _ in range(s['num_rpl1']):
                _ref_pic_list_struct(r, s['long_term_refs'], s['poc_bits'], forbid)
        try:
            _sps_after_lists(r, s)
        except ValueError:
            continue
        if (not (s['weighted_pred'] or s['weighted_bipred'])) == forbid and r.at_trailing_bits():
            return s
    raise ValueError('SPS does not end on rbsp_trailing_bits: not the syntax this reader follows')


def _sps_after_lists(r, s):
    s['dual_tree'] = r.flag()
    r.u(2)                                                           # log2_ctu_size_minus5
    r.ue()                                                           # log2_min_luma_coding_block_size_minus2
    s['partition_override_enabled'] = r.flag()
    r.ue(); r.ue()
    depth_inter, depth_intra = r.ue(), r.ue()
    if depth_intra:
        r.ue(); r.ue()
    if depth_inter:
        r.ue(); r.ue()
    if s['dual_tree']:
        r.ue()
        if r.ue():
            r.ue(); r.ue()
    r.flag()                                                         # sps_max_luma_transform_size_64_flag
    if s['chroma_format_idc'] != 0:
        same = r.flag()
        for _ in range(1 if same else 3):
            for _ in range(r.ue() + 1):
                r.ue(); r.ue()
    s['weighted_pred'], s['weighted_bipred'] = r.flag(), r.flag()
    s['sao'], s['alf'] = r.flag(), r.flag()
    s['transform_skip'] = r.flag()
    if s['transform_skip']:
        r.flag()                                                     # sps_bdpcm_enabled_flag
    s['joint_cbcr'] = r.flag()
    if r.flag():                                                     # sps_ref_wraparound_enabled_flag
        r.ue()
    s['temporal_mvp'] = r.flag()
    if s['temporal_mvp']:
        r.flag()
    r.flag()                                                         # amvr
    bdof = r.flag()
    dmvr = r.flag()
    mmvd = r.flag()
    if r.flag() and s['chroma_format_idc'] == 1:                     # lm_chroma_enabled_flag
        r.flag()
    if r.flag():                                                     # mts_enabled_flag
        r.flag(); r.flag()
    r.flag(); r.flag()                                               # lfnst, smvd
    if r.flag():                                                     # affine
        r.flag(); r.flag(); r.flag()
    r.flag()                                                         # gbi
    if s['chroma_format_idc'] == 3:
        r.flag()
    s['ibc'] = r.flag()
    r.flag()                                                         # mhintra
    if mmvd:
        r.flag()
    if bdof or dmvr:
        r.flag()
    r.flag(); r.flag()                                               # triangle, mip
    if r.flag():                                                     # sbt
        r.flag()
    s['lmcs'] = r.flag()
    r.flag()                                                         # isp
    if r.flag():                                                     # ladf
        n = r.u(2) + 2
        r.se()
        for _ in range(1, n):
            r.se(); r.ue()
    if r.flag():
        raise NotImplementedError('scaling lists')
    if r.flag():                                                     # timing_info_present_flag
        r.u(32); r.u(32)
        if r.flag():
            raise NotImplementedError('HRD parameters')
    if r.flag():
        raise NotImplementedError('VUI')
    if r.flag():                                                     # sps_extension_present_flag
        ext = [r.flag() for _ in range(8)]
        if any(ext[1:]):
            raise NotImplementedError('SPS extension')
        if ext[0]:
            r.u(9)


def parse_pps(rbsp, sps_by_id):
    """EL/VLCWriter.cpp:208-491 -> dict of the fields the slice header depends on."""
    r, p = BitReader(rbsp), {}
    p['pps_id'], p['sps_id'] = r.ue(), r.ue()
    r.ue(); r.ue()
    if r.flag():
        r.ue(); r.ue(); r.ue(); r.ue()
    p['output_flag_present'] = r.flag()
    p['num_extra_slice_header_bits'] = r.u(3)
    p['cabac_init_present'] = r.flag()
    r.ue(); r.ue()
    p['rpl1_idx_present'] = r.flag()
    p['init_qp_minus26'] = r.se()
    r.flag()                                                         # constrained_intra_pred_flag
    if r.flag():                                                     # cu_qp_delta_enabled_flag
        r.ue()
    r.se(); r.se(); r.se()
    p['slice_chroma_qp_offsets_present'] = r.flag()
    r.flag(); r.flag(); r.flag()                                     # weighted pred / bipred, transquant bypass
    if not r.flag():
        raise NotImplementedError('several tiles per picture')
    p['rect_slice'], p['single_brick_per_slice'], p['num_bricks'], p['lf_across_slices'] = 1, 1, 1, 0
    p['signalled_slice_id'] = r.flag()
    if p['signalled_slice_id']:
        p['slice_id_len'] = r.ue() + 1
        r.u(p['slice_id_len'])                                       # one slice
    p['entropy_coding_sync'] = r.flag()
    p['deblocking_control_present'] = r.flag()
    p['deblocking_override_enabled'], p['deblocking_disabled'] = 0, 0
    if p['deblocking_control_present']:
        p['deblocking_override_enabled'] = r.flag()
        p['deblocking_disabled'] = r.flag()
        if not p['deblocking_disabled']:
            r.se(); r.se()
    if r.flag():
        raise NotImplementedError('virtual boundaries')
    if r.flag():
        raise NotImplementedError('scaling lists')
    r.ue()
    p['slice_header_extension_present'] = r.flag()
    p['chroma_qp_offset_list_enabled'] = 0
    if r.flag():
        ext = [r.flag() for _ in range(8)]
        if any(ext[1:]):
            raise NotImplementedError('PPS extension')
        if ext[0]:                                                   # :455-481
            if p['sps_id'] in sps_by_id and sps_by_id[p['sps_id']]['transform_skip']:
                r.ue()                                               # log2_max_transform_skip_block_size_minus2
            r.flag()                                                 # cross_component_prediction_enabled_flag
            p['chroma_qp_offset_list_enabled'] = r.flag()
            if p['chroma_qp_offset_list_enabled']:
                r.ue()
                for _ in range(r.ue() + 1):
                    r.se(); r.se(); r.se()
            r.ue(); r.ue()                                           # log2_sao_offset_scale_luma / chroma
    if not r.at_trailing_bits():
        raise ValueError('PPS does not end on rbsp_trailing_bits: not the syntax this reader follows')
    if p['sps_id'] not in sps_by_id:
        raise ValueError('PPS refers to an SPS that was not seen')
    return p


def parse_intra_slice_header(rbsp, nal_type, sps, pps_by_id):
    """EL/VLCWriter.cpp:1167-1699 for an intra slice of a single-tile picture.  Returns a dict with the bit positions the gather edits:
    `poc_pos` (first bit of slice_pic_order_cnt_lsb), `after_poc`, `after_rpl` (== after_poc for an IDR picture without sps_idr_rpl_present_flag),
    `header_end` (the alignment one-bit), `data_start` (first byte of the slice data), and the values read on the way."""
    r, h = BitReader(rbsp), {}
    rap = nal_type in (NAL_IDR_W_RADL, NAL_IDR_N_LP, NAL_CRA)
    idr = nal_type in (NAL_IDR_W_RADL, NAL_IDR_N_LP)
    if rap:
        h['no_output_of_prior_pics_pos'] = r.pos
        h['no_output_of_prior_pics'] = r.flag()
    pps = pps_by_id[r.ue()]
    r.u(pps['slice_id_len'] if pps['signalled_slice_id'] else 1)     # slice_address: a single-tile picture has rect_slice_flag 1 (:404), one slice -> 1 bit
    for _ in range(pps['num_extra_slice_header_bits']):
        r.flag()
    h['slice_type'] = r.ue()
    if h['slice_type'] != I_SLICE:
        raise NotImplementedError('inter slice')
    if pps['output_flag_present']:
        r.flag()
    h['poc_pos'] = r.pos
    h['poc_lsb'] = r.u(sps['poc_bits'])
    h['after_poc'] = r.pos
    if not idr or sps['idr_rpl_present']:
        forbid = not (sps['weighted_pred'] or sps['weighted_bipred'])
        idx0 = None
        sps_flag0 = r.flag() if sps['num_rpl0'] > 0 else 0
        if sps_flag0:
            idx0 = r.u((sps['num_rpl0'] - 1).bit_length()) if sps['num_rpl0'] > 1 else 0
        else:
            _ref_pic_list_struct(r, sps['long_term_refs'], sps['poc_bits'], forbid)
        if not pps['rpl1_idx_present']:
            if not sps_flag0:
                _ref_pic_list_struct(r, sps['long_term_refs'], sps['poc_bits'], forbid)
        else:
            sps_flag1 = r.flag() if sps['num_rpl1'] > 0 else 0
            if sps_flag1:
                if sps['num_rpl1'] > 1:
                    r.u((sps['num_rpl1'] - 1).bit_length())
            else:
                _ref_pic_list_struct(r, sps['long_term_refs'], sps['poc_bits'], forbid)
        if sps['temporal_mvp']:
            r.flag()
    h['after_rpl'] = r.pos
    sao_luma = sao_chroma = 0
    chroma = sps['chroma_format_idc'] != 0
    if sps['sao']:
        sao_luma = r.flag()
        if chroma:
            sao_chroma = r.flag()
    if sps['alf']:
        h['alf'] = r.flag()
        if h['alf']:
            h['alf_aps_pos'] = []
            n = r.u(3)
            for _ in range(n):
                h['alf_aps_pos'].append(r.pos)
                r.u(3)
            idc = r.u(2) if chroma else 0
            if idc:
                h['alf_aps_chroma_pos'] = r.pos
                r.u(3)
    if not r.flag():                                                 # dep_quant_enabled_flag
        r.flag()
    if sps['partition_override_enabled']:
        if r.flag():
            r.ue()
            if r.ue():
                r.ue(); r.ue()
            if sps['dual_tree']:
                r.ue()
                if r.ue():
                    r.ue(); r.ue()
    if sps['ibc']:
        r.ue()
    if sps['joint_cbcr']:
        r.flag()
    h['slice_qp_delta'] = r.se()
    if pps['slice_chroma_qp_offsets_present']:
        if chroma:
            r.se(); r.se()
            if sps['joint_cbcr']:
                r.se()
    if pps['chroma_qp_offset_list_enabled']:
        r.flag()
    dbf_disabled = pps['deblocking_disabled']
    if pps['deblocking_control_present']:
        override = r.flag() if pps['deblocking_override_enabled'] else 0
        if override:
            dbf_disabled = r.flag()
            if not dbf_disabled:
                r.se(); r.se()
    if pps['lf_across_slices'] and (sao_luma or sao_chroma or not dbf_disabled):
        r.flag()
    if sps['lmcs']:
        if r.flag():
            h['lmcs_aps_pos'] = r.pos
            r.u(2)
            if chroma:
                r.flag()
    if pps['slice_header_extension_present']:
        if r.ue():
            raise NotImplementedError('slice header extension')
    if pps['entropy_coding_sync']:
        raise NotImplementedError('entry points (wavefronts)')
    h['header_end'] = r.pos
    if r.u(1) != 1:
        raise ValueError('slice header does not end on the alignment bit: not the syntax this reader follows')
    while not r.byte_aligned():
        if r.u(1):
            raise ValueError('slice header does not end on the alignment bits: not the syntax this reader follows')
    h['data_start'] = r.pos >> 3
    return h
