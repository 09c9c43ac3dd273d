/*
 * TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the reference's intra cost-evaluation path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (libvvc_intra_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  Every function here is checked against records captured from the
 * unmodified reference encoder (oracle/ref_trace_hooks.cpp -> tests/golden/ fixtures) by
 * tests/test_oracle_golden.py.
 */
#ifndef VVC_ORACLE_H
#define VVC_ORACLE_H
#include <stdint.h>
#include "../include/vvc_intra_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* IntraPrediction::m_ipaParam as derived by initPredIntraParams (CL/IntraPrediction.cpp:487-618) */
typedef struct orc_ipa {
  int is_ver, mrl, ref_filter, interp, pdpc, angle, inv_angle, ang_scale;
} orc_ipa;

/* a1/a2: reference lines.  top has 2w+1+mrl samples, left has 2h+1+mrl; top[0] == left[0]. */
void orc_ref_fill(const int16_t* reco, int stride, int x, int y, int w, int h, int mrl, int bd,
                  int avail_al, int n_above, int n_above_right, int n_left, int n_below_left,
                  int16_t* top, int16_t* left);
/* a3 */
void orc_ref_filter(const int16_t* top, const int16_t* left, int w, int h, int mrl, int16_t* ftop, int16_t* fleft);
/* a4 */
void orc_ipa_init(int w, int h, int mode, int mrl, int is_mip, orc_ipa* p);
/* a5: regular prediction from the lines selected by the caller (filtered iff p->ref_filter) */
void orc_pred_regular(const int16_t* top, const int16_t* left, int w, int h, int bd, int mode,
                      const orc_ipa* p, int16_t* pred);
/* a6: MIP prediction from unfiltered line-0 references */
int  orc_mip_num_modes(int w, int h);
void orc_pred_mip(const int16_t* top, const int16_t* left, int w, int h, int bd, int mode, int16_t* pred);
/* a7/a8 */
uint64_t orc_sad (const int16_t* org, int org_stride, const int16_t* cur, int cur_stride, int w, int h);
uint64_t orc_satd(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride, int w, int h);
/* a10 */
uint64_t orc_mode_bits(const vvcb_rates* r, const uint8_t mpm[6], int w, int h, int mrl_allowed,
                       int mip_enabled, int is_mip, int mrl, int mode);
void orc_intra_mpms(int left_dir, int above_dir, uint8_t mpm[6], int* num_cand);
/* a9: the whole RMD of one visit, references fetched from the reco plane.
 * pred_out (optional) receives VVCB_NUM_SLOTS blocks of w*h samples. */
void orc_rmd_visit(const int16_t* orig, int orig_stride, const int16_t* reco, int reco_stride,
                   int bd, int ctu_size, const vvcb_rmd_visit* v, vvcb_rmd_result* out, vvcb_rmd_detail* det,
                   int16_t* pred_out);
/* batch with an OpenMP-free pthread pool is overkill for a checker: plain loop */
void orc_rmd_batch(const int16_t* orig, int orig_stride, const int16_t* reco, int reco_stride,
                   int bd, int ctu_size, const vvcb_rmd_visit* v, int n, vvcb_rmd_result* out, vvcb_rmd_detail* det);
uint64_t orc_fnv1a(const int16_t* p, int n);

/* ---- TU coding (vvc_oracle_tr.c) ---- */
void orc_tr_types(int mts_idx, int* hor, int* ver);
void orc_fwd_transform(const int16_t* resi, int stride, int w, int h, int bd, int mts_idx, int32_t* coeff);
void orc_fwd_transform_ex(const int16_t* resi, int stride, int w, int h, int bd, int mts_idx, int lfnst_idx, int32_t* coeff);
void orc_transform_skip(const int16_t* resi, int stride, int w, int h, int bd, int32_t* coeff);
int  orc_abs_sum_for_preselection(const int32_t* coeff, int w, int h, int mts_idx);
void orc_mts_preselect(const int* sums, int n, int w, int h, int max_cand, uint8_t* selected);
int  orc_quant_scalar(const int32_t* coeff, int w, int h, int bd, int per, int rem, int is_ts, int32_t* level);
void orc_dequant(const int32_t* level, int w, int h, int bd, int per, int rem, int is_ts, int32_t* coeff);
void orc_inv_transform(const int32_t* coeff, int w, int h, int bd, int mts_idx, int16_t* resi, int stride);
void orc_inv_transform_skip(const int32_t* coeff, int w, int h, int bd, int16_t* resi, int stride);
uint64_t orc_reconstruct_sse(const int16_t* org, int org_stride, const int16_t* pred, const int16_t* resi, int w, int h, int bd, int16_t* reco);

/* LFNST (vvc_oracle_tr.c), in place on a dense w*h coefficient block; intra_mode = PU::getFinalIntraMode (planar for MIP) */
void orc_fwd_lfnst(int32_t* coeff, int w, int h, int intra_mode, int lfnst_idx);
void orc_inv_lfnst(int32_t* coeff, int w, int h, int intra_mode, int lfnst_idx);

/* ---- dependent quantisation (vvc_oracle_dq.c); qp = QpParam::Qp of the block ---- */
int  orc_dep_quant(const int32_t* coeff, int w, int h, int bd, int mts_idx, int lfnst_idx, int qp, double lambda,
                   const vvcb_dq_rates* rates, int cbf_delta_bits, int32_t* level);
void orc_dep_dequant(const int32_t* level, int w, int h, int bd, int qp, int32_t* coeff);

/* ---- RDOQ of transform-skip blocks (vvc_oracle_rdoq.c); qp = QpParam::Qp(true) ---- */
int  orc_rdoq_ts(const int32_t* coeff, int w, int h, int bd, int qp, double lambda, const vvcb_dq_rates* rates, int32_t* level);

/* ---- residual rate estimation (vvc_oracle_rate.c): fractional bits of CABACWriter::residual_coding( tu, COMPONENT_Y ) ---- */
uint64_t orc_residual_bits(const int32_t* level, int w, int h, int mts_idx, int ts_allowed, int mts_allowed, int dep_quant,
                           const vvcb_ctx_states* states);

/* ---- texture measures (vvc_oracle_feat.c) ---- */
void orc_ctu_hads_islice(const int16_t* orig, int stride, int pic_w, int pic_h, int ctu, int32_t* out);
void orc_features(const int16_t* orig, int stride, const vvcb_feat_job* job, vvcb_feat_result* out);
void orc_features_batch(const int16_t* orig, int stride, const vvcb_feat_job* jobs, int n, vvcb_feat_result* out);

#ifdef __cplusplus
}
#endif
#endif
