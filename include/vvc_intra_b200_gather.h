/*
 * vvc_intra_b200 gather -- the final step of the frame-parallel encode: per-picture bitstreams -> one bitstream.
 *
 * All-intra pictures are independent, so the pictures of a sequence are encoded by independent encoder processes (one per picture,
 * `-f 1 --FrameSkip=k`, picture k on GPU k mod N behind that GPU's broker) and gathered afterwards.  Host code only: no CUDA device
 * is needed or touched by these two calls.  What they replace in the reference: APP/Parcat/parcat.cpp (main :418-446,
 * filter_segment :247-384), with the slice-header knowledge of EL/VLCWriter.cpp:1167-1699 (codeSliceHeader), :736-1130 (codeSPS),
 * :208-491 (codePPS), EL/NALwrite.cpp:47-140 (NAL header, emulation prevention) and EL/EncGOP.cpp:4279-4290 (header alignment).
 *
 * Errors: VVCB_ERR_ARG for bad arguments, unreadable files and streams that are not what the call expects; VVCB_ERR_STATE for syntax
 * the reader deliberately does not follow (several tiles, scaling lists, VUI / HRD, inter slices, long-term reference pictures);
 * the message goes to `err` (NUL-terminated, at most err_len bytes) when err is not NULL.
 */
#ifndef VVC_INTRA_B200_GATHER_H
#define VVC_INTRA_B200_GATHER_H
#include "vvc_intra_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* The bit-exact gather.  paths[0..n-1]: one-picture bitstreams (parameter sets + one IDR slice, POC 0) in picture order.  Writes to
 * out_path the bitstream the sequential encoder (`EncoderApp -f n`) writes for the sequence, byte for byte: the slice NAL unit of
 * picture k >= 1 becomes NAL_UNIT_CODED_SLICE_CRA with slice_pic_order_cnt_lsb = k and the reference-picture-list syntax of a
 * non-IDR slice (EL/VLCWriter.cpp:1238-1335: list 0 of the SPS, GOP position 0, EL/EncLib.cpp:1632; slice_temporal_mvp_enabled_flag
 * as EL/EncGOP.cpp:2226), the rest of the header shifted and re-aligned, emulation prevention redone.  rewrite_param_sets follows the
 * encoder's ReWriteParamSets (1 in the shipped configuration: VPS / SPS / PPS ahead of every picture; 0: only ahead of the first).
 * bytes_written may be NULL.                                                                                                        */
int vvcb_gather_sequential(const char* const* paths, int n, const char* out_path, int rewrite_param_sets, uint64_t* bytes_written,
                           char* err, int err_len);

/* The reference's Parcat on the same inputs, byte for byte (same argument order: segments, then the output): random-access segments
 * that overlap by their IDR picture; of segments 2.. the parameter sets / access unit delimiters ahead of the IDR picture, the IDR
 * picture itself and the suffix SEI behind it are dropped, every other slice has the number of non-IDR pictures gathered so far
 * added to its 8-bit slice_pic_order_cnt_lsb (including the tool's habit of keeping the old low bit, parcat.cpp:337-339).
 * pictures_renumbered may be NULL.                                                                                                 */
int vvcb_gather_parcat(const char* const* paths, int n, const char* out_path, int* pictures_renumbered, char* err, int err_len);

#ifdef __cplusplus
}
#endif
#endif
