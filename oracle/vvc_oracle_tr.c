/*
 * TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the TU-coding arithmetic of the reference:
 * forward / inverse separable transforms, transform skip, scalar quantisation, dequantisation,
 * reconstruction and SSE (SURVEY.md 8a rows a11, a12, a14 and the scalar part of a13).
 * See vvc_oracle.h for who may use it.  Checked against the reference's own outputs by
 * tests/test_oracle_tu.py (records 'S', 'Q', 'I' of oracle/ref_trace_hooks.cpp).
 *
 * The reference evaluates the 1-D transforms with partial butterflies / fast DST-VII factorizations
 * (CL/TrQuant_EMT.cpp); they are exact integer evaluations of the matrix products below (one rounding
 * per output, no intermediate rounding, no int32 overflow for 8..10-bit residuals), so the plain
 * products are bit-identical -- which the golden records confirm.
 */
#include <stdlib.h>
#include <string.h>
#include "vvc_oracle.h"
#include "../vvc_intra_b200/csrc/vvc_rom_tables.h"

static int ilog2(int v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; }

/* kernel of a transform type (0 DCT-II, 1 DCT-VIII, 2 DST-VII: CL/TypeDef.h TransType) and size */
static const int16_t* kernel_of(int type, int n)
{
  switch (type * 8 + ilog2(n)) {
    case 0 * 8 + 2: return kDct2_4;  case 0 * 8 + 3: return kDct2_8;  case 0 * 8 + 4: return kDct2_16;
    case 0 * 8 + 5: return kDct2_32; case 0 * 8 + 6: return kDct2_64;
    case 1 * 8 + 2: return kDct8_4;  case 1 * 8 + 3: return kDct8_8;  case 1 * 8 + 4: return kDct8_16; case 1 * 8 + 5: return kDct8_32;
    case 2 * 8 + 2: return kDst7_4;  case 2 * 8 + 3: return kDst7_8;  case 2 * 8 + 4: return kDst7_16; case 2 * 8 + 5: return kDst7_32;
  }
  return NULL;
}

/* TrQuant::getTrTypes for explicit intra MTS (CL/TrQuant.cpp:752-831): mts_idx 0 DCT2xDCT2, 2..5 = DST7/DCT8 pairs */
void orc_tr_types(int mts_idx, int* hor, int* ver)
{
  *hor = 0; *ver = 0;
  if (mts_idx > 1) {
    *hor = ((mts_idx - 2) & 1) ? 1 : 2;
    *ver = ((mts_idx - 2) >> 1) ? 1 : 2;
  }
}

static void skips(int w, int h, int hor, int ver, int* skipW, int* skipH)
{
  *skipW = (hor != 0 && w == 32) ? 16 : (w > 32 ? w - 32 : 0);     /* CL/TrQuant.cpp:853-854 */
  *skipH = (ver != 0 && h == 32) ? 16 : (h > 32 ? h - 32 : 0);
}

/* a12  TrQuant::xT (CL/TrQuant.cpp:835-915): rows first, then columns, high frequencies zeroed */
void orc_fwd_transform(const int16_t* resi, int stride, int w, int h, int bd, int mts_idx, int32_t* coeff)
{
  int hor, ver, skipW, skipH, j, k, n, l;
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * w * h);
  const int shift1 = ilog2(w) + bd + 6 - 15, shift2 = ilog2(h) + 6;
  const int add1 = shift1 > 0 ? 1 << (shift1 - 1) : 0, add2 = 1 << (shift2 - 1);
  const int16_t *mh, *mv;
  orc_tr_types(mts_idx, &hor, &ver);
  skips(w, h, hor, ver, &skipW, &skipH);
  mh = kernel_of(hor, w); mv = kernel_of(ver, h);
  memset(coeff, 0, sizeof(int32_t) * w * h);
  for (j = 0; j < h; j++)
    for (k = 0; k < w - skipW; k++) {
      int acc = 0;
      for (n = 0; n < w; n++) acc += mh[k * w + n] * resi[j * stride + n];
      tmp[k * h + j] = (acc + add1) >> shift1;
    }
  for (k = 0; k < w - skipW; k++)
    for (l = 0; l < h - skipH; l++) {
      int acc = 0;
      for (j = 0; j < h; j++) acc += mv[l * h + j] * tmp[k * h + j];
      coeff[l * w + k] = (acc + add2) >> shift2;
    }
  free(tmp);
}

static int transform_shift(int bd, int w, int h) { return 15 - bd - ((ilog2(w) + ilog2(h)) >> 1); }   /* getTransformShift */

/* TrQuant::xTransformSkip (CL/TrQuant.cpp:1394-1438) */
void orc_transform_skip(const int16_t* resi, int stride, int w, int h, int bd, int32_t* coeff)
{
  const int sh = transform_shift(bd, w, h);
  int x, y;
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) coeff[y * w + x] = (int32_t)resi[y * stride + x] << sh;
}

/* Sum of |coeff| as TrQuant::transformNxN(trModes) ranks candidates (CL/TrQuant.cpp:1090-1103) */
int orc_abs_sum_for_preselection(const int32_t* coeff, int w, int h, int mts_idx)
{
  int s = 0, i;
  double scale = 1.0;
  for (i = 0; i < w * h; i++) s += abs(coeff[i]);
  if (mts_idx == 1 && ((ilog2(w) + ilog2(h)) & 1)) scale = 1.0 / 1.414213562;
  return (int)(s * scale);
}

/* MTS pre-selection (CL/TrQuant.cpp:1112-1123): sums[] in candidate order (DCT2 first, transform skip second if present) */
void orc_mts_preselect(const int* sums, int n, int w, int h, int max_cand, uint8_t* selected)
{
  static const double facBB[5] = { 1.2, 1.3, 1.3, 1.4, 1.5 };
  const double fac = facBB[ilog2(w > h ? w : h) - 2];
  const double thr = fac * sums[0], thrTS = sums[0];
  int i, tests = 0;
  for (i = 0; i < n; i++) {
    const int t = sums[i] <= (i == 1 ? thrTS : thr) && tests <= max_cand;
    selected[i] = (uint8_t)t;
    tests += t;
  }
}

/* a13 (scalar part)  Quant::quant (CL/Quant.cpp:994-1089) without scaling lists / sign hiding, intra (IRAP) rounding */
int orc_quant_scalar(const int32_t* coeff, int w, int h, int bd, int per, int rem, int is_ts, int32_t* level)
{
  const int sqrtAdj = !is_ts && ((ilog2(w) + ilog2(h)) & 1);
  const int scale = kQuantScales[(sqrtAdj ? 6 : 0) + rem];
  const int qbits = 14 + per + transform_shift(bd, w, h) + (sqrtAdj ? -1 : 0);
  const int64_t add = (int64_t)171 << (qbits - 9);
  int i, absSum = 0;
  for (i = 0; i < w * h; i++) {
    const int64_t t = (int64_t)abs(coeff[i]) * scale;
    const int32_t mag = (int32_t)((t + add) >> qbits);
    int32_t q = coeff[i] < 0 ? -mag : mag;
    absSum += mag;
    if (q < -32768) q = -32768;
    if (q > 32767) q = 32767;
    level[i] = q;
  }
  return absSum;
}

/* Quant::dequant (CL/Quant.cpp:423-540), flat scaling */
void orc_dequant(const int32_t* level, int w, int h, int bd, int per, int rem, int is_ts, int32_t* coeff)
{
  const int sqrtAdj = !is_ts && ((ilog2(w) + ilog2(h)) & 1);
  const int scale = kInvQuantScales[(sqrtAdj ? 6 : 0) + rem];
  const int rightShift = 6 - (transform_shift(bd, w, h) + (sqrtAdj ? -1 : 0) + per);
  int tgt = 32 + rightShift - 7, i;
  int inMin, inMax;
  if (tgt > 16) tgt = 16;
  inMin = -(1 << (tgt - 1)); inMax = (1 << (tgt - 1)) - 1;
  for (i = 0; i < w * h; i++) {
    int q = level[i], c;
    if (q < inMin) q = inMin;
    if (q > inMax) q = inMax;
    if (rightShift > 0) c = (q * scale + (1 << (rightShift - 1))) >> rightShift;
    else                c = (int)((unsigned)(q * scale) << (-rightShift));
    if (c < -32768) c = -32768;
    if (c > 32767) c = 32767;
    coeff[i] = c;
  }
}

/* a14  TrQuant::xIT (CL/TrQuant.cpp:917-993): columns first (shift 7), then rows, clipping after each stage */
void orc_inv_transform(const int32_t* coeff, int w, int h, int bd, int mts_idx, int16_t* resi, int stride)
{
  int hor, ver, skipW, skipH, j, k, n, x, y;
  int32_t* tmp = (int32_t*)calloc((size_t)w * h, sizeof(int32_t));
  const int shift1 = 7, shift2 = 20 - bd;
  const int16_t *mh, *mv;
  orc_tr_types(mts_idx, &hor, &ver);
  skips(w, h, hor, ver, &skipW, &skipH);
  mh = kernel_of(hor, w); mv = kernel_of(ver, h);
  for (j = 0; j < w - skipW; j++)
    for (n = 0; n < h; n++) {
      int acc = 0, v;
      for (k = 0; k < h - skipH; k++) acc += mv[k * h + n] * coeff[k * w + j];
      v = (acc + (1 << (shift1 - 1))) >> shift1;
      tmp[j * h + n] = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
    }
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) {
      int acc = 0, v;
      for (k = 0; k < w - skipW; k++) acc += mh[k * w + x] * tmp[k * h + y];
      v = (acc + (1 << (shift2 - 1))) >> shift2;
      v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
      resi[y * stride + x] = (int16_t)v;
    }
  free(tmp);
}

/* TrQuant::xITransformSkip (CL/TrQuant.cpp:996-1041) */
void orc_inv_transform_skip(const int32_t* coeff, int w, int h, int bd, int16_t* resi, int stride)
{
  const int sh = transform_shift(bd, w, h);
  const int off = sh == 0 ? 0 : 1 << (sh - 1);
  int x, y;
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) resi[y * stride + x] = (int16_t)((coeff[y * w + x] + off) >> sh);
}

/* PelBuf::reconstruct + RdCost::xGetSSE (CL/Buffer.cpp, CL/RdCost.cpp:1739) */
uint64_t orc_reconstruct_sse(const int16_t* org, int org_stride, const int16_t* pred, const int16_t* resi, int w, int h, int bd, int16_t* reco)
{
  uint64_t sse = 0;
  const int maxv = (1 << bd) - 1;
  int x, y;
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) {
      int r = pred[y * w + x] + resi[y * w + x], d;
      r = r < 0 ? 0 : (r > maxv ? maxv : r);
      reco[y * w + x] = (int16_t)r;
      d = org[y * org_stride + x] - r;
      sse += (uint64_t)(d * d);
    }
  return sse;
}
