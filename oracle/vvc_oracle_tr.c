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

void orc_fwd_transform(const int16_t* resi, int stride, int w, int h, int bd, int mts_idx, int32_t* coeff)
{
  orc_fwd_transform_ex(resi, stride, w, h, bd, mts_idx, 0, coeff);
}

/* a12  TrQuant::xT (CL/TrQuant.cpp:835-915): rows first, then columns, high frequencies zeroed; with LFNST only the
 * top-left 4x4 / 8x8 of primary coefficients is produced (JVET_O0094, :853-867) */
void orc_fwd_transform_ex(const int16_t* resi, int stride, int w, int h, int bd, int mts_idx, int lfnst_idx, int32_t* coeff)
{
  int hor, ver, skipW, skipH, j, k, n, l;
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * w * h);
  const int shift1 = ilog2(w) + bd + 6 - 15, shift2 = ilog2(h) + 6;
  const int add1 = shift1 > 0 ? 1 << (shift1 - 1) : 0, add2 = 1 << (shift2 - 1);
  const int16_t *mh, *mv;
  orc_tr_types(mts_idx, &hor, &ver);
  skips(w, h, hor, ver, &skipW, &skipH);
  if (lfnst_idx) {
    if ((w == 4 && h > 4) || (w > 4 && h == 4)) { skipW = w - 4; skipH = h - 4; }
    else if (w >= 8 && h >= 8) { skipW = w - 8; skipH = h - 8; }
  }
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

/* ---- LFNST: TrQuant::xFwdLfnst / xInvLfnst, fwdLfnstNxN / invLfnstNxN (CL/TrQuant.cpp:239-560) ---- */

/* PU::getWideAngIntraMode (CL/UnitTools.cpp:963-989) then TrQuant::getLFNSTIntraMode (CL/TrQuant.cpp:288-306): 0..94 */
static int lfnst_intra_mode(int w, int h, int dirMode)
{
  static const int modeShift[] = { 0, 6, 10, 12, 14, 15 };
  int predMode = dirMode;
  if (dirMode >= 2) {
    const int delta = abs(ilog2(w) - ilog2(h));
    if (w > h && dirMode < 2 + modeShift[delta]) predMode += 66 - 1;
    else if (h > w && predMode > 66 - modeShift[delta]) predMode -= 66 + 1;
  }
  if (predMode < 0) return predMode + (28 >> 1) + 67;          /* NUM_EXT_LUMA_MODE 28, NUM_LUMA_MODE 67 */
  if (predMode >= 67) return predMode + (28 >> 1);
  return predMode;
}

static int lfnst_transpose(int m) { return (m >= 67 && m >= 67 + (28 >> 1)) || (m < 67 && m > 34); }   /* getTransposeFlag :308-312 */

/* scan of the LFNST region: the top-left 8x8 in 4x4 groups (g_coefTopLeftDiagScan8x8, CL/Rom.cpp:369-392) when w,h >= 8,
 * else the block's own grouped scan, whose first 16 entries are the diagonal scan of the top-left 4x4 */
static void lfnst_scan(int w, int whge3, int* idx)
{
  int gx[4], gy[4], ix[16], iy[16], n = 0, d, y, g, i;
  for (d = 0; d <= 6; d++) for (y = d < 3 ? d : 3; y >= 0 && d - y <= 3; y--) { ix[n] = d - y; iy[n] = y; n++; }
  gx[0] = 0; gy[0] = 0; gx[1] = 0; gy[1] = 1; gx[2] = 1; gy[2] = 0; gx[3] = 1; gy[3] = 1;      /* diagonal order of a 2x2 grid */
  for (g = 0; g < (whge3 ? 4 : 1); g++)
    for (i = 0; i < 16; i++) idx[g * 16 + i] = (gy[g] * 4 + iy[i]) * w + gx[g] * 4 + ix[i];
}

void orc_fwd_lfnst(int32_t* coeff, int w, int h, int intra_mode, int lfnst_idx)
{
  const int whge3 = w >= 8 && h >= 8, sb = whge3 ? 8 : 4, trSize = whge3 ? 48 : 16;
  const int zeroOut = ((w == 4 && h == 4) || (w == 8 && h == 8)) ? 8 : 16;
  const int m = lfnst_intra_mode(w, h, intra_mode), tr = lfnst_transpose(m), set = kLfnstLut[m];
  const int8_t* mat = whge3 ? kLfnst8x8 + (set * 2 + lfnst_idx - 1) * 16 * 48 : kLfnst4x4 + (set * 2 + lfnst_idx - 1) * 16 * 16;
  int in[48], out[48], scan[64], x, y, j, i, n = 0;
  if (!lfnst_idx) return;
  if (tr) {                                     /* column-major gathering, the bottom-right 4x4 of an 8x8 is left out */
    if (sb == 4) { for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) in[y + 4 * x] = coeff[y * w + x]; }
    else for (y = 0; y < 8; y++) {
      for (x = 0; x < 4; x++) in[y + 8 * x] = coeff[y * w + x];
      if (y < 4) for (x = 4; x < 8; x++) in[32 + y + 4 * (x - 4)] = coeff[y * w + x];
    }
  } else {
    for (y = 0; y < sb; y++) { const int len = y < 4 ? sb : 4; for (x = 0; x < len; x++) in[n++] = coeff[y * w + x]; }
  }
  for (j = 0; j < trSize; j++) out[j] = 0;
  for (j = 0; j < zeroOut; j++) {
    int acc = 0;
    for (i = 0; i < trSize; i++) acc += in[i] * mat[j * trSize + i];
    out[j] = (acc + 64) >> 7;
  }
  lfnst_scan(w, whge3, scan);
  for (j = 0; j < (sb == 4 ? 16 : 48); j++) coeff[scan[j]] = out[j];
}

void orc_inv_lfnst(int32_t* coeff, int w, int h, int intra_mode, int lfnst_idx)
{
  const int whge3 = w >= 8 && h >= 8, sb = whge3 ? 8 : 4, trSize = whge3 ? 48 : 16;
  const int zeroOut = ((w == 4 && h == 4) || (w == 8 && h == 8)) ? 8 : 16;
  const int m = lfnst_intra_mode(w, h, intra_mode), tr = lfnst_transpose(m), set = kLfnstLut[m];
  const int8_t* mat = whge3 ? kLfnst8x8 + (set * 2 + lfnst_idx - 1) * 16 * 48 : kLfnst4x4 + (set * 2 + lfnst_idx - 1) * 16 * 16;
  int in[16], out[48], scan[64], x, y, j, i, n = 0;
  if (!lfnst_idx) return;
  lfnst_scan(w, whge3, scan);
  for (i = 0; i < 16; i++) in[i] = coeff[scan[i]];
  for (j = 0; j < trSize; j++) {
    int acc = 0, v;
    for (i = 0; i < zeroOut; i++) acc += in[i] * mat[i * trSize + j];
    v = (acc + 64) >> 7;
    out[j] = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
  }
  if (tr) {
    if (sb == 4) { for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) coeff[y * w + x] = out[y + 4 * x]; }
    else for (y = 0; y < 8; y++) {
      for (x = 0; x < 4; x++) coeff[y * w + x] = out[y + 8 * x];
      if (y < 4) for (x = 4; x < 8; x++) coeff[y * w + x] = out[32 + y + 4 * (x - 4)];
    }
  } else {
    for (y = 0; y < sb; y++) { const int len = y < 4 ? sb : 4; for (x = 0; x < len; x++) coeff[y * w + x] = out[n++]; }
  }
}
