/*
 * TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the reference's intra rough-mode-decision
 * path, in plain C.  See vvc_oracle.h for the rules on who may use it.
 *
 * Each function cites the reference code it restates (CL/ = VVC_project/source/Lib/CommonLib/,
 * EL/ = VVC_project/source/Lib/EncoderLib/).  Compile with -ffp-contract=off: the costs are IEEE
 * doubles and must be evaluated in the reference's operation order without fused multiply-adds.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "vvc_oracle.h"
#include "../vvc_intra_b200/csrc/vvc_rom_tables.h"

#define MAXN 64

static int ilog2(int v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static int iclip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------------
 * a1/a2  IntraPrediction::xFillReferenceSamples, CL/IntraPrediction.cpp:1215-1468.
 *
 * The reference walks 4-sample "units" from the bottom-left corner up the left column, over the
 * above-left corner and along the top row, copying available units and padding the others from the
 * nearest earlier unit on that walk (or, before the first available unit, from that unit's first
 * sample).  Restated here as one linear walk over samples.
 * ---------------------------------------------------------------------------------------------- */
void orc_ref_fill(const int16_t* reco, int stride, int x, int y, int w, int h, int mrl, int bd,
                  int avail_al, int n_above, int n_above_right, int n_left, int n_below_left,
                  int16_t* top, int16_t* left)
{
  const int nTop = 2 * w + 1 + mrl, nLeft = 2 * h + 1 + mrl;
  /* walk: left[nLeft-1] ... left[1], corner top[0], top[1] ... top[nTop-1] */
  const int len = (nLeft - 1) + nTop;
  int16_t val[4 * MAXN + 16];
  uint8_t ok[4 * MAXN + 16];
  int i;

  for (i = 0; i < len; i++) {
    int isLeft = i < nLeft - 1;
    int k = isLeft ? (nLeft - 1 - i) : (i - (nLeft - 1));   /* index into left[] or top[] */
    int px, py, avail;
    if (isLeft) {
      px = x - 1 - mrl; py = y - 1 - mrl + k;
      if (k <= mrl) avail = avail_al;
      else {
        int unit = (k - 1 - mrl) >> 2;
        avail = unit < (h >> 2) ? unit < n_left : (unit - (h >> 2)) < n_below_left;
      }
    } else {
      px = x - 1 - mrl + k; py = y - 1 - mrl;
      if (k <= mrl) avail = avail_al;
      else {
        int unit = (k - 1 - mrl) >> 2;
        avail = unit < (w >> 2) ? unit < n_above : (unit - (w >> 2)) < n_above_right;
      }
    }
    ok[i] = (uint8_t)avail;
    val[i] = avail ? reco[py * stride + px] : 0;
  }

  {
    int first = 0;
    while (first < len && !ok[first]) first++;
    if (first == len) {                       /* :1279-1284 nothing available: mid-grey */
      for (i = 0; i < len; i++) val[i] = (int16_t)(1 << (bd - 1));
    } else {
      for (i = 0; i < first; i++) val[i] = val[first];          /* :1358-1403 */
      for (i = first + 1; i < len; i++) if (!ok[i]) val[i] = val[i - 1];   /* :1405-1458 */
    }
  }
  for (i = 0; i < len; i++) {
    if (i < nLeft - 1) left[nLeft - 1 - i] = val[i];
    else               top[i - (nLeft - 1)] = val[i];
  }
  left[0] = top[0];
}

/* a3  IntraPrediction::xFilterReferenceSamples, CL/IntraPrediction.cpp:1470-1522: [1 2 1]/4 along the
 * L-shaped line, end points untouched. */
void orc_ref_filter(const int16_t* top, const int16_t* left, int w, int h, int mrl, int16_t* ftop, int16_t* fleft)
{
  const int nTop = 2 * w + 1 + mrl, nLeft = 2 * h + 1 + mrl;
  int i;
  ftop[0] = (int16_t)((left[1] + 2 * top[0] + top[1] + 2) >> 2);
  for (i = 1; i < nTop - 1; i++) ftop[i] = (int16_t)((top[i - 1] + 2 * top[i] + top[i + 1] + 2) >> 2);
  ftop[nTop - 1] = top[nTop - 1];
  fleft[0] = ftop[0];
  for (i = 1; i < nLeft - 1; i++) {
    int prev = i == 1 ? top[0] : left[i - 1];
    fleft[i] = (int16_t)((prev + 2 * left[i] + left[i + 1] + 2) >> 2);
  }
  fleft[nLeft - 1] = left[nLeft - 1];
}

/* ------------------------------------------------------------------------------------------------
 * a4  IntraPrediction::initPredIntraParams (CL/IntraPrediction.cpp:487-618) and getWideAngle (:287)
 * for luma CUs without ISP/BDPCM.
 * ---------------------------------------------------------------------------------------------- */
static const int kAngTable[32] = { 0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29,
                                   32, 35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512, 1024 };
static const int kInvAngTable[32] = { 0, 16384, 8192, 5461, 4096, 2731, 2048, 1638, 1365, 1170, 1024, 910, 819, 712, 630, 565,
                                      512, 468, 420, 364, 321, 287, 256, 224, 191, 161, 128, 96, 64, 48, 32, 16 };
static const int kIntraFilterThr[8] = { 24, 24, 24, 14, 2, 0, 0, 0 };   /* m_aucIntraFilter :58-74 */

static int wide_angle(int w, int h, int mode)
{
  static const int shiftTab[6] = { 0, 6, 10, 12, 14, 15 };
  if (mode > 1 && mode <= 66) {
    int d = abs(ilog2(w) - ilog2(h));
    if (w > h && mode < 2 + shiftTab[d]) mode += 65;
    else if (h > w && mode > 66 - shiftTab[d]) mode -= 65;
  }
  return mode;
}

void orc_ipa_init(int w, int h, int mode, int mrl, int is_mip, orc_ipa* p)
{
  const int predMode = is_mip ? mode : wide_angle(w, h, mode);
  int angMode, absAng = 0;
  memset(p, 0, sizeof(*p));
  p->is_ver = predMode >= 34;
  p->mrl = mrl;
  p->pdpc = (w >= 4 && h >= 4) && mrl == 0;
  angMode = p->is_ver ? predMode - 50 : -(predMode - 18);
  if (is_mip) { /* PU::getFinalIntraMode maps MIP to planar for the angle set-up, but no filtering (:563) */
    p->is_ver = 0; p->pdpc = mrl == 0; return;
  }
  if (mode > 1 && mode < 67) {
    const int a = abs(angMode);
    absAng = kAngTable[a];
    p->inv_angle = kInvAngTable[a];
    p->angle = angMode < 0 ? -absAng : absAng;
    if (angMode < 0) p->pdpc = 0;
    else if (angMode > 0) {
      const int side = p->is_ver ? h : w;
      p->ang_scale = imin(2, ilog2(side) - (ilog2(3 * p->inv_angle - 2) - 8));
      p->pdpc = p->pdpc && p->ang_scale >= 0;
    }
  }
  if (mrl || mode == 1) return;                      /* :559-575 */
  if (mode == 0) { p->ref_filter = w * h > 32; return; }   /* :580-583 */
  {
    const int diff = imin(abs(predMode - 18), abs(predMode - 50));
    const int log2Size = (ilog2(w) + ilog2(h)) >> 1;
    if (diff > kIntraFilterThr[log2Size]) {
      const int integerSlope = (absAng & 31) == 0;
      p->ref_filter = integerSlope;
      p->interp = !integerSlope;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * a5  predIntraAng (CL/IntraPrediction.cpp:316-398): planar :426, DC :248/:480, angular :633-935,
 * PDPC :354-378 / :771-782 / :844-865.
 * ---------------------------------------------------------------------------------------------- */
static void pred_planar(const int16_t* top, const int16_t* left, int w, int h, int16_t* pred)
{
  const int lw = ilog2(w), lh = ilog2(h);
  const int tr = top[w + 1], bl = left[h + 1];
  int x, y;
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) {
      /* closed form of the running sums at :451-478 */
      int hor = (left[y + 1] << lw) + (x + 1) * (tr - left[y + 1]);
      int ver = (top[x + 1] << lh) + (y + 1) * (bl - top[x + 1]);
      pred[y * w + x] = (int16_t)(((hor << lh) + (ver << lw) + (1 << (lw + lh))) >> (1 + lw + lh));
    }
}

static int dc_value(const int16_t* top, const int16_t* left, int w, int h, int mrl)
{
  int sum = 0, i;
  const int denom = w == h ? 2 * w : imax(w, h);
  if (w >= h) for (i = 0; i < w; i++) sum += top[mrl + 1 + i];
  if (w <= h) for (i = 0; i < h; i++) sum += left[mrl + 1 + i];
  return (sum + (denom >> 1)) >> ilog2(denom);
}

static void pdpc_planar_dc(const int16_t* top, const int16_t* left, int w, int h, int16_t* pred)
{
  const int scale = (ilog2(w) + ilog2(h) - 2) >> 2;
  int x, y;
  for (y = 0; y < h; y++) {
    const int wT = 32 >> imin(31, (y << 1) >> scale);
    for (x = 0; x < w; x++) {
      const int wL = 32 >> imin(31, (x << 1) >> scale);
      const int v = pred[y * w + x];
      pred[y * w + x] = (int16_t)(v + ((wL * (left[y + 1] - v) + wT * (top[x + 1] - v) + 32) >> 6));
    }
  }
}

static void pred_angular(const int16_t* top, const int16_t* left, int w, int h, int bd,
                         const orc_ipa* p, int16_t* pred)
{
  /* Work in the "main/side" frame: for horizontal modes the roles of top/left and of x/y swap and
   * the result is transposed on output (:747-754, :924-934). */
  const int16_t* mainSrc = p->is_ver ? top : left;
  const int16_t* sideSrc = p->is_ver ? left : top;
  const int mw = p->is_ver ? w : h;      /* extent along the main reference */
  const int mh = p->is_ver ? h : w;      /* extent along the side reference = number of "rows" */
  const int mrl = p->mrl, angle = p->angle, maxv = (1 << bd) - 1;
  int16_t mainBuf[3 * MAXN + 32], sideBuf[3 * MAXN + 32];
  int16_t* refMain = mainBuf + MAXN;     /* room for negative indices */
  int16_t* refSide = sideBuf;
  int i, r, c;

  if (angle < 0) {                        /* :654-673 */
    for (i = 0; i <= mw + 1 + mrl; i++) refMain[i] = mainSrc[i];
    for (i = 0; i <= mh + 1 + mrl; i++) refSide[i] = sideSrc[i];
    for (i = -mh; i <= -1; i++) refMain[i] = refSide[imin((-i * p->inv_angle + 256) >> 9, mh)];
  } else {                                /* :702-726 */
    const int mainLen = 2 * mw, sideLen = 2 * mh;
    const int s = imax(0, ilog2(mw) - ilog2(mh));
    const int ext = (mrl << s) + 2;
    for (i = 0; i <= mainLen + mrl; i++) refMain[i] = mainSrc[i];
    for (i = 0; i <= sideLen + mrl; i++) refSide[i] = sideSrc[i];
    for (i = 1; i <= ext; i++) refMain[mainLen + mrl + i] = refMain[mainLen + mrl];
  }
  refMain += mrl;
  refSide += mrl;

  for (r = 0; r < mh; r++) {
    int line[MAXN];
    if (angle == 0) {                     /* :762-786 */
      for (c = 0; c < mw; c++) line[c] = refMain[c + 1];
      if (p->pdpc) {
        const int scale = (ilog2(mw) + ilog2(mh) - 2) >> 2;
        const int lim = imin(3 << scale, mw);
        for (c = 0; c < lim; c++) {
          const int wL = 32 >> ((2 * c) >> scale);
          line[c] = iclip(line[c] + ((wL * (refSide[1 + r] - refMain[0]) + 32) >> 6), 0, maxv);
        }
      }
    } else {                              /* :789-866 */
      const int pos = angle * (r + 1 + mrl);
      const int dInt = pos >> 5, dFrac = pos & 31;
      if (abs(angle) & 31) {
        const int8_t* f = (p->interp ? kIntraGaussFilter : kIntraCubicFilter) + 4 * dFrac;
        for (c = 0; c < mw; c++) {
          const int16_t* q = refMain + dInt + c;
          line[c] = iclip((f[0] * q[0] + f[1] * q[1] + f[2] * q[2] + f[3] * q[3] + 32) >> 6, 0, maxv);
        }
      } else {
        for (c = 0; c < mw; c++) line[c] = refMain[c + dInt + 1];
      }
      if (p->pdpc) {
        const int scale = p->ang_scale;
        const int lim = imin(3 << scale, mw);
        int acc = 256;
        for (c = 0; c < lim; c++) {
          int wL, l;
          acc += p->inv_angle;
          wL = 32 >> ((2 * c) >> scale);
          l = refSide[r + (acc >> 9) + 1];
          line[c] = line[c] + ((wL * (l - line[c]) + 32) >> 6);
        }
      }
    }
    for (c = 0; c < mw; c++) {
      if (p->is_ver) pred[r * w + c] = (int16_t)line[c];
      else           pred[c * w + r] = (int16_t)line[c];
    }
  }
}

void orc_pred_regular(const int16_t* top, const int16_t* left, int w, int h, int bd, int mode,
                      const orc_ipa* p, int16_t* pred)
{
  if (mode == 0) {
    pred_planar(top, left, w, h, pred);
    if (p->pdpc) pdpc_planar_dc(top, left, w, h, pred);
  } else if (mode == 1) {
    const int dc = dc_value(top, left, w, h, p->mrl);
    int i;
    for (i = 0; i < w * h; i++) pred[i] = (int16_t)dc;
    if (p->pdpc) pdpc_planar_dc(top, left, w, h, pred);
  } else {
    pred_angular(top, left, w, h, bd, p, pred);
  }
}

/* ------------------------------------------------------------------------------------------------
 * a6  Matrix-based intra prediction: CL/MatrixIntraPrediction.cpp (prepareInputForPred :71,
 * predBlock :211, computeReducedPred :637, predictionUpsampling :469), getNumModesMip
 * CL/UnitTools.cpp:4688.
 * ---------------------------------------------------------------------------------------------- */
int orc_mip_num_modes(int w, int h)
{
  if (w > 4 * h || h > 4 * w) return 0;
  if (w == 4 && h == 4) return 35;
  if (w <= 8 && h <= 8) return 19;
  return 11;
}

static void mip_downsample(const int16_t* src, int srcLen, int dstLen, int* dst)
{
  const int f = srcLen / dstLen, lf = ilog2(f);
  int i, k;
  for (i = 0; i < dstLen; i++) {
    if (f == 1) { dst[i] = src[i]; continue; }
    {
      int s = 0;
      for (k = 0; k < f; k++) s += src[i * f + k];
      dst[i] = (s + (1 << (lf - 1))) >> lf;
    }
  }
}

void orc_pred_mip(const int16_t* top, const int16_t* left, int w, int h, int bd, int mode, int16_t* pred)
{
  const int numModes = orc_mip_num_modes(w, h);
  const int transpose = mode > numModes / 2;
  const int widx = transpose ? mode - numModes / 2 : mode;
  const int small = w <= 8 && h <= 8;                 /* 4x4 or 8x8 matrix family, full first column */
  const int bsz = (w > 4 || h > 4) ? 4 : 2;           /* reduced boundary size per side */
  const int redW = small ? 4 : imin(w, 8), redH = small ? 4 : imin(h, 8);
  const int upH = w / redW, upV = h / redH;
  const int inSize = 2 * bsz;
  const uint8_t* mat; int shift, offs, cols, grid;
  int bT[4], bL[4], in[8], red[64];                   /* red[] holds the reduced prediction, logical [y][x] */
  int i, xx, yy;

  if (w == 4 && h == 4) { mat = kMipMatrix4x4 + widx * 16 * 4; shift = kMipShift4x4[widx]; offs = kMipOffset4x4[widx]; cols = 4; grid = 4; }
  else if (small)       { mat = kMipMatrix8x8 + widx * 16 * 8; shift = kMipShift8x8[widx]; offs = kMipOffset8x8[widx]; cols = 8; grid = 4; }
  else                  { mat = kMipMatrix16x16 + widx * 64 * 7; shift = kMipShift16x16[widx]; offs = kMipOffset16x16[widx]; cols = 7; grid = 8; }

  mip_downsample(top + 1, w, bsz, bT);
  mip_downsample(left + 1, h, bsz, bL);
  for (i = 0; i < bsz; i++) {
    in[i]       = transpose ? bL[i] : bT[i];
    in[bsz + i] = transpose ? bT[i] : bL[i];
  }
  {
    const int inOff = in[0];
    int sum = 0, off;
    in[0] = small ? inOff - (1 << (bd - 1)) : 0;
    for (i = 1; i < inSize; i++) in[i] -= inOff;
    for (i = 0; i < inSize; i++) sum += in[i];
    off = (1 << (shift - 1)) - offs * sum;
    {
      /* matrix product in the (possibly transposed) frame; rows/columns of the 8x8 output grid are
       * skipped for 4xN / Nx4 blocks (leaveHorOut / leaveVerOut, :230-235) */
      int lho = (w == 4 && h >= 16), lvo = (h == 4 && w >= 16);
      const int iw = transpose ? redH : redW, ih = transpose ? redW : redH;
      if (transpose) { int t = lho; lho = lvo; lvo = t; }
      for (yy = 0; yy < ih; yy++)
        for (xx = 0; xx < iw; xx++) {
          const int row = (lvo ? 2 * yy : yy) * (small ? iw : grid) + (lho ? 2 * xx : xx);
          const uint8_t* wgt = mat + row * cols;
          int acc = 0, v;
          if (small) for (i = 0; i < inSize; i++) acc += in[i] * wgt[i];
          else       for (i = 1; i < inSize; i++) acc += in[i] * wgt[i - 1];
          v = iclip(((acc + off) >> shift) + inOff, 0, (1 << bd) - 1);
          if (transpose) red[xx * redW + yy] = v; else red[yy * redW + xx] = v;
        }
    }
  }

  if (upH == 1 && upV == 1) {
    for (i = 0; i < w * h; i++) pred[i] = (int16_t)red[i];
    return;
  }
  {
    /* Separable linear interpolation, shorter side first (:469-567); each pass anchors on the
     * original boundary samples.  tmp holds the first pass (values at reduced positions only). */
    int tmp[MAXN * MAXN];
    const int lH = ilog2(upH), lV = ilog2(upV);
    int x, y, k;
    if (h > w) {
      /* horizontal first on the rows that carry reduced samples: y = upV*(ry+1)-1 */
      for (yy = 0; yy < redH; yy++) {
        const int row = upV * (yy + 1) - 1;
        for (xx = 0; xx < redW; xx++) {
          const int before = xx == 0 ? left[1 + row] : red[yy * redW + xx - 1];
          const int behind = red[yy * redW + xx];
          for (k = 1; k <= upH; k++) {
            const int v = upH == 1 ? behind : ((upH - k) * before + k * behind + (1 << (lH - 1))) >> lH;
            tmp[row * w + xx * upH + k - 1] = v;
          }
        }
      }
      for (x = 0; x < w; x++)
        for (yy = 0; yy < redH; yy++) {
          const int before = yy == 0 ? top[1 + x] : tmp[(upV * yy - 1) * w + x];
          const int behind = tmp[(upV * (yy + 1) - 1) * w + x];
          for (k = 1; k <= upV; k++)
            pred[(yy * upV + k - 1) * w + x] = (int16_t)(((upV - k) * before + k * behind + (1 << (lV - 1))) >> lV);
        }
    } else {
      /* vertical first on the columns that carry reduced samples: x = upH*(rx+1)-1 */
      for (xx = 0; xx < redW; xx++) {
        const int col = upH * (xx + 1) - 1;
        for (yy = 0; yy < redH; yy++) {
          const int before = yy == 0 ? top[1 + col] : red[(yy - 1) * redW + xx];
          const int behind = red[yy * redW + xx];
          for (k = 1; k <= upV; k++) {
            const int v = upV == 1 ? behind : ((upV - k) * before + k * behind + (1 << (lV - 1))) >> lV;
            tmp[(yy * upV + k - 1) * w + col] = v;
          }
        }
      }
      for (y = 0; y < h; y++)
        for (xx = 0; xx < redW; xx++) {
          const int before = xx == 0 ? left[1 + y] : tmp[y * w + upH * xx - 1];
          const int behind = tmp[y * w + upH * (xx + 1) - 1];
          for (k = 1; k <= upH; k++)
            pred[y * w + xx * upH + k - 1] = (int16_t)(upH == 1 ? behind : ((upH - k) * before + k * behind + (1 << (lH - 1))) >> lH);
        }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * a7  RdCost::xGetSAD (CL/RdCost.cpp:449)   a8  RdCost::xGetHADs (:2746-2861) and its tiles.
 * The tiles are plain 2-D Walsh-Hadamard transforms; the sum of absolute coefficients does not depend
 * on the butterfly order, only the per-tile normalisation matters.
 * ---------------------------------------------------------------------------------------------- */
uint64_t orc_sad(const int16_t* org, int os, const int16_t* cur, int cs, int w, int h)
{
  uint64_t s = 0; int x, y;
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) s += (uint64_t)abs(org[y * os + x] - cur[y * cs + x]);
  return s;
}

static int wht_abs_sum(const int16_t* org, int os, const int16_t* cur, int cs, int tw, int th)
{
  int d[16 * 16];
  int x, y, len, i, s = 0;
  for (y = 0; y < th; y++) for (x = 0; x < tw; x++) d[y * tw + x] = org[y * os + x] - cur[y * cs + x];
  for (y = 0; y < th; y++)
    for (len = 1; len < tw; len <<= 1)
      for (i = 0; i < tw; i += 2 * len)
        for (x = i; x < i + len; x++) {
          int a = d[y * tw + x], b = d[y * tw + x + len];
          d[y * tw + x] = a + b; d[y * tw + x + len] = a - b;
        }
  for (x = 0; x < tw; x++)
    for (len = 1; len < th; len <<= 1)
      for (i = 0; i < th; i += 2 * len)
        for (y = i; y < i + len; y++) {
          int a = d[y * tw + x], b = d[(y + len) * tw + x];
          d[y * tw + x] = a + b; d[(y + len) * tw + x] = a - b;
        }
  for (i = 0; i < tw * th; i++) s += abs(d[i]);
  return s;
}

uint64_t orc_satd(const int16_t* org, int os, const int16_t* cur, int cs, int w, int h)
{
  int tw, th, x, y;
  uint64_t sum = 0;
  if      (w > h && (h & 7) == 0 && (w & 15) == 0) { tw = 16; th = 8; }
  else if (w < h && (w & 7) == 0 && (h & 15) == 0) { tw = 8; th = 16; }
  else if (w > h && (h & 3) == 0 && (w & 7) == 0)  { tw = 8; th = 4; }
  else if (w < h && (w & 3) == 0 && (h & 7) == 0)  { tw = 4; th = 8; }
  else if ((h & 7) == 0 && (w & 7) == 0)           { tw = 8; th = 8; }
  else                                             { tw = 4; th = 4; }
  for (y = 0; y < h; y += th)
    for (x = 0; x < w; x += tw) {
      const int s = wht_abs_sum(org + y * os + x, os, cur + y * cs + x, cs, tw, th);
      if (tw == 4 && th == 4)      sum += (uint64_t)((s + 1) >> 1);                    /* :2209 */
      else if (tw == 8 && th == 8) sum += (uint64_t)((s + 2) >> 2);                    /* :2306 */
      else if (tw * th == 128)     sum += (uint64_t)(int)(s / sqrt(16.0 * 8) * 2);     /* :2452, :2589 */
      else                         sum += (uint64_t)(int)(s / sqrt(4.0 * 8) * 2);      /* :2662, :2741 */
    }
  return sum;
}

/* ------------------------------------------------------------------------------------------------
 * a10  bits of CABACWriter::intra_luma_pred_mode (EL/CABACWriter.cpp:1762-1845) on the estimator:
 * mip_flag :4741, mip_pred_mode :4781, extend_ref_line :1566, isp_mode :3944, xWriteTruncBinCode :1543.
 * ---------------------------------------------------------------------------------------------- */
static int trunc_bin_len(int symbol, int numSymbols)
{
  const int thresh = ilog2(numSymbols);
  const int b = numSymbols - (1 << thresh);
  return symbol < (1 << thresh) - b ? thresh : thresh + 1;
}

uint64_t orc_mode_bits(const vvcb_rates* r, const uint8_t mpm[6], int w, int h, int mrl_allowed,
                       int mip_enabled, int is_mip, int mrl, int mode)
{
  const uint64_t EP = 1u << 15;
  uint64_t bits = 0;
  int idx = 6, i;
  if (mip_enabled && w <= 64 && h <= 64 && orc_mip_num_modes(w, h)) bits += r->mip_flag[is_mip ? 1 : 0];
  if (is_mip) return bits + EP * (uint64_t)trunc_bin_len(mode, orc_mip_num_modes(w, h));
  if (mrl_allowed) {
    bits += r->mrl_bin0[mrl != 0];
    if (mrl != 0) bits += r->mrl_bin1[mrl != 1];
  }
  if (mrl == 0 && ilog2(w) + ilog2(h) > 4 && w <= 64 && h <= 64) bits += r->isp_bin0_0;   /* CU::canUseISP */
  for (i = 0; i < 6; i++) if (mpm[i] == mode) { idx = i; break; }
  if (mrl == 0) bits += r->mpm_flag[idx < 6];
  if (idx < 6) {
    if (mrl == 0) bits += r->planar_flag[idx > 0];
    if (idx > 0) bits += EP * (uint64_t)imin(idx, 4);
  } else {
    uint8_t s[6]; int j, rem = mode;
    memcpy(s, mpm, 6);
    for (i = 1; i < 6; i++) { uint8_t k = s[i]; for (j = i; j > 0 && s[j - 1] > k; j--) s[j] = s[j - 1]; s[j] = k; }
    for (i = 5; i >= 0; i--) if (rem > s[i]) rem--;
    bits += EP * (uint64_t)trunc_bin_len(rem, 61);
  }
  return bits;
}

/* PU::getIntraMPMs, CL/UnitTools.cpp:507-640, from the two neighbour directions. */
void orc_intra_mpms(int L, int A, uint8_t mpm[6], int* num_cand)
{
  const int offset = 61, mod = 64;
  mpm[0] = 0; mpm[1] = 1; mpm[2] = 50; mpm[3] = 18; mpm[4] = 46; mpm[5] = 54;
  if (L == A) {
    *num_cand = 1;
    if (L > 1) {
      mpm[1] = (uint8_t)L;
      mpm[2] = (uint8_t)(((L + offset) % mod) + 2);
      mpm[3] = (uint8_t)(((L - 1) % mod) + 2);
      mpm[4] = (uint8_t)(((L + offset - 1) % mod) + 2);
      mpm[5] = (uint8_t)((L % mod) + 2);
    }
  } else {
    *num_cand = 2;
    if (L > 1 && A > 1) {
      const int mx = imax(L, A), mn = imin(L, A);
      mpm[1] = (uint8_t)L; mpm[2] = (uint8_t)A;
      if (mx - mn == 1)       { mpm[3] = (uint8_t)(((mn + offset) % mod) + 2); mpm[4] = (uint8_t)(((mx - 1) % mod) + 2); mpm[5] = (uint8_t)(((mn + offset - 1) % mod) + 2); }
      else if (mx - mn >= 62) { mpm[3] = (uint8_t)(((mn - 1) % mod) + 2); mpm[4] = (uint8_t)(((mx + offset) % mod) + 2); mpm[5] = (uint8_t)((mn % mod) + 2); }
      else if (mx - mn == 2)  { mpm[3] = (uint8_t)(((mn - 1) % mod) + 2); mpm[4] = (uint8_t)(((mn + offset) % mod) + 2); mpm[5] = (uint8_t)(((mx - 1) % mod) + 2); }
      else                    { mpm[3] = (uint8_t)(((mn + offset) % mod) + 2); mpm[4] = (uint8_t)(((mn - 1) % mod) + 2); mpm[5] = (uint8_t)(((mx + offset) % mod) + 2); }
    } else if (L + A >= 2) {
      const int m = imax(L, A);
      mpm[1] = (uint8_t)m;
      mpm[2] = (uint8_t)(((m + offset) % mod) + 2);
      mpm[3] = (uint8_t)(((m - 1) % mod) + 2);
      mpm[4] = (uint8_t)(((m + offset - 1) % mod) + 2);
      mpm[5] = (uint8_t)((m % mod) + 2);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * a9  candidate lists: updateCandList (CL/UnitTools.h:261-307) is a stable bounded insertion with a
 * strict '<'; reduceHadCandList is EL/IntraSearch.cpp:4333-4405; the schedule is :430-802.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { vvcb_mode m[VVCB_MAX_LIST + 4]; double c[VVCB_MAX_LIST + 4]; int n; } cand_list;

static void cand_push(cand_list* L, vvcb_mode m, double cost, int cap)
{
  const int live = L->n < cap ? L->n : cap;
  int pos = live, i;
  while (pos > 0 && cost < L->c[pos - 1]) pos--;
  if (L->n >= cap) {
    if (pos == live) return;
    for (i = live - 1; i > pos; i--) { L->m[i] = L->m[i - 1]; L->c[i] = L->c[i - 1]; }
  } else {
    for (i = L->n; i > pos; i--) { L->m[i] = L->m[i - 1]; L->c[i] = L->c[i - 1]; }
    L->n++;
  }
  L->m[pos] = m; L->c[pos] = cost;
}

static int same_mode(vvcb_mode a, vvcb_mode b) { return a.mip == b.mip && a.mrl == b.mrl && a.mode == b.mode; }
static vvcb_mode mk_mode(int mip, int mrl, int mode) { vvcb_mode m; m.mip = (uint8_t)mip; m.mrl = (uint8_t)mrl; m.mode = (uint8_t)mode; m.pad = 0; return m; }

static void copy_list(const cand_list* L, int32_t* n, vvcb_mode* m, double* c, int cap)
{
  int i;
  *n = L->n;
  for (i = 0; i < L->n && i < cap; i++) { m[i] = L->m[i]; if (c) c[i] = L->c[i]; }
}

static const uint8_t kFastModes[6][6] = {   /* g_aucIntraModeNumFast_UseMPM_2D, CL/Rom.cpp:536 */
  { 3, 3, 3, 3, 2, 2 }, { 3, 3, 3, 3, 3, 2 }, { 3, 3, 3, 3, 3, 2 }, { 3, 3, 3, 3, 3, 2 }, { 2, 3, 3, 3, 3, 2 }, { 2, 2, 2, 2, 2, 3 } };

void orc_rmd_visit(const int16_t* orig, int orig_stride, const int16_t* reco, int reco_stride,
                   int bd, int ctu_size, const vvcb_rmd_visit* v, vvcb_rmd_result* out, vvcb_rmd_detail* det,
                   int16_t* pred_out)
{
  vvcb_rmd_detail detLocal;
  const int w = 1 << v->log2w, h = 1 << v->log2h;
  const int16_t* org = orig + v->y * orig_stride + v->x;
  const int mipEnabled = !(v->flags & VVCB_VISIT_NO_MIP);
  const int numMip = (mipEnabled && w <= 64 && h <= 64) ? orc_mip_num_modes(w, h) : 0;
  const int testMip = numMip > 0;
  const int mrlAllowed = !(v->flags & VVCB_VISIT_NO_MRL) && (v->y & (ctu_size - 1)) != 0;
  int16_t top[3][2 * MAXN + 8], left[3][2 * MAXN + 8], ftop[2 * MAXN + 8], fleft[2 * MAXN + 8];
  int16_t pred[MAXN * MAXN];
  static const int kMrl[3] = { 0, 1, 3 };
  double cost[VVCB_NUM_SLOTS], dist[VVCB_NUM_SLOTS];
  uint8_t checked[VVCB_NUM_LUMA_MODE];
  cand_list rd, had, parent;
  int K = kFastModes[v->log2w - 2][v->log2h - 2];
  int numHad, slot, i, li, m;

  if (!det) det = &detLocal;
  memset(out, 0, sizeof(*out));
  memset(det, 0, sizeof(*det));
  for (i = 0; i < VVCB_NUM_SLOTS; i++) { det->sad[i] = VVCB_SAT_NONE; det->satd[i] = VVCB_SAT_NONE; }
  for (li = 0; li < 3; li++)
    orc_ref_fill(reco, reco_stride, v->x, v->y, w, h, kMrl[li], bd, v->avail_al, v->n_above, v->n_above_right,
                 v->n_left, v->n_below_left, top[li], left[li]);
  orc_ref_filter(top[0], left[0], w, h, 0, ftop, fleft);

  /* evaluate every slot */
  for (slot = 0; slot < VVCB_NUM_SLOTS; slot++) {
    int isMip = slot >= VVCB_SLOT_MIP, mrl = 0, mode = slot;
    uint64_t sad, satd, mn, bits;
    orc_ipa p;
    li = 0;
    if (isMip) { mode = slot - VVCB_SLOT_MIP; if (mode >= numMip) continue; }
    else if (slot >= VVCB_SLOT_MRL1) {
      li = slot >= VVCB_SLOT_MRL3 ? 2 : 1;
      mrl = kMrl[li];
      mode = v->mpm[1 + (slot - VVCB_SLOT_MRL1) % 5];
      if (!mrlAllowed) continue;
    }
    if (isMip) orc_pred_mip(top[0], left[0], w, h, bd, mode, pred);
    else {
      orc_ipa_init(w, h, mode, mrl, 0, &p);
      if (p.ref_filter) orc_pred_regular(ftop, fleft, w, h, bd, mode, &p, pred);
      else              orc_pred_regular(top[li], left[li], w, h, bd, mode, &p, pred);
    }
    if (pred_out) memcpy(pred_out + (size_t)slot * w * h, pred, sizeof(int16_t) * w * h);
    sad = orc_sad(org, orig_stride, pred, w, w, h);
    satd = orc_satd(org, orig_stride, pred, w, w, h);
    det->sad[slot] = (uint32_t)sad; det->satd[slot] = (uint32_t)satd;
    mn = sad * 2 < satd ? sad * 2 : satd;                                   /* :515 */
    bits = orc_mode_bits(&v->rates, v->mpm, w, h, mrlAllowed, mipEnabled, isMip, mrl, mode);
    dist[slot] = (double)mn;
    cost[slot] = (double)mn + (double)bits * v->sqrt_lambda;                /* :526 */
  }

  /* replay the reference's insertion schedule */
  if (testMip) K += imax(K, ilog2(imin(w, h)) - 1);                         /* :472 (FastMIP) */
  numHad = testMip ? 6 : 3;
  rd.n = had.n = 0;
  memset(checked, 0, sizeof(checked));
  for (m = 0; m < VVCB_NUM_LUMA_MODE; m++) {                                /* :489-532 */
    if (m > 1 && (m & 1)) continue;
    checked[m] = 1;
    cand_push(&rd, mk_mode(0, 0, m), cost[m], K);
    cand_push(&had, mk_mode(0, 0, m), dist[m], numHad);
  }
  parent = rd;
  for (i = 0; i < K; i++) {                                                 /* :577-623 */
    const int pm = parent.m[i].mode;
    int d;
    if (pm > 2 && pm < 66)
      for (d = -1; d <= 1; d += 2) {
        m = pm + d;
        if (checked[m]) continue;
        cand_push(&rd, mk_mode(0, 0, m), cost[m], K);
        cand_push(&had, mk_mode(0, 0, m), dist[m], numHad);
        checked[m] = 1;
      }
  }
  if (mrlAllowed)                                                           /* :635-681 */
    for (li = 1; li < 3; li++)
      for (i = 1; i < 6; i++) {
        slot = (li == 1 ? VVCB_SLOT_MRL1 : VVCB_SLOT_MRL3) + i - 1;
        cand_push(&rd, mk_mode(0, kMrl[li], v->mpm[i]), cost[slot], K);
        cand_push(&had, mk_mode(0, kMrl[li], v->mpm[i]), dist[slot], numHad);
      }
  copy_list(&rd, &det->n_reg, det->reg_mode, det->reg_cost, VVCB_MAX_LIST);
  copy_list(&had, &det->n_reg_had, det->reg_had_mode, det->reg_had_cost, VVCB_MAX_HAD_LIST);

  if (testMip) {                                                            /* :704-751 */
    double mipCost[35];
    cand_list tmp;
    const double thr = 1.0 + 1.4 / sqrt((double)(w * h));
    const int maxPerType = K >> 1;
    int keepOne, numConv = 0, numMipKept = 0, idx;
    for (m = 0; m < numMip; m++) {
      slot = VVCB_SLOT_MIP + m;
      mipCost[m] = cost[slot];
      cand_push(&rd, mk_mode(1, 0, m), cost[slot], K + 1);
      cand_push(&had, mk_mode(1, 0, m), 0.8 * dist[slot], numHad);
    }
    /* reduceHadCandList :4333 */
    tmp.n = 0;
    keepOne = rd.n > K;
    for (idx = 0; idx < rd.n - (keepOne ? 0 : 1); idx++) {
      int add;
      if (!rd.m[idx].mip) { add = numConv < 3; numConv += add; }
      else {
        add = numMipKept < maxPerType || rd.c[idx] < thr * rd.c[0] || keepOne;
        keepOne = 0;
        numMipKept += add;
      }
      if (add) { tmp.m[tmp.n] = rd.m[idx]; tmp.c[tmp.n] = rd.c[idx]; tmp.n++; }
    }
    if (w > 8 && h > 8) {
      const int off = numMip / 2;
      cand_list srt; int base = tmp.n;
      srt.n = 0;
      for (m = 3; m <= 5; m++) {
        const int cm = m + (mipCost[m + off] < mipCost[m] ? off : 0);
        cand_push(&srt, mk_mode(1, 0, cm), mipCost[cm], 3);
      }
      for (idx = 0; idx < 3; idx++) {
        int inc = 0;
        for (i = 0; i < base; i++) if (same_mode(tmp.m[i], srt.m[idx])) { inc = 1; break; }
        if (!inc) { tmp.m[tmp.n] = srt.m[idx]; tmp.c[tmp.n] = 0; tmp.n++; break; }   /* fastMip: first one only */
      }
    }
    rd = tmp;
    K = rd.n;
  }
  copy_list(&rd, &out->n_rd, out->rd_mode, out->rd_cost, VVCB_MAX_LIST);
  copy_list(&had, &out->n_had, out->had_mode, out->had_cost, VVCB_MAX_HAD_LIST);

  for (i = 0; i < v->num_mpm_cand; i++) {                                   /* :777-802 */
    const vvcb_mode mp = mk_mode(0, 0, v->mpm[i]);
    int inc = 0, j;
    for (j = 0; j < K; j++) inc |= same_mode(mp, rd.m[j]);
    if (!inc) { rd.m[rd.n] = mp; rd.c[rd.n] = 0; rd.n++; K++; }
  }
  copy_list(&rd, &out->n_final, out->final_mode, NULL, VVCB_MAX_LIST);
}

void orc_rmd_batch(const int16_t* orig, int orig_stride, const int16_t* reco, int reco_stride,
                   int bd, int ctu_size, const vvcb_rmd_visit* v, int n, vvcb_rmd_result* out, vvcb_rmd_detail* det)
{
  int i;
  for (i = 0; i < n; i++) orc_rmd_visit(orig, orig_stride, reco, reco_stride, bd, ctu_size, v + i, out + i, det ? det + i : NULL, NULL);
}

uint64_t orc_fnv1a(const int16_t* p, int n)
{
  uint64_t hsh = 1469598103934665603ull;
  const uint8_t* q = (const uint8_t*)p;
  int i;
  for (i = 0; i < 2 * n; i++) { hsh ^= q[i]; hsh *= 1099511628211ull; }
  return hsh;
}
