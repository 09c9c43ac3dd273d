/*
 * TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the reference's rate-distortion optimised quantisation of
 * transform-skip blocks: QuantRDOQ::xRateDistOptQuantTS (CL/QuantRDOQ.cpp:1243-1485) with xGetCodedLevelTSPred
 * (:2000-2065), xGetICRateTS (:2067-2150), xGetErrScaleCoeff (:383-393) and the transform-skip context derivations of
 * CoeffCodingContext (CL/ContextModelling.h:197-365, CL/ContextModelling.cpp:114-134); luma, no BDPCM,
 * JVET_O0122_TS_SIGN_LEVEL on as shipped.  See vvc_oracle.h for who may use it.
 *
 * Parity status: PINNED ('T' records of oracle/ref_trace_hooks.cpp, tests/test_oracle_rdoq.py).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "vvc_oracle.h"
#include "../vvc_intra_b200/csrc/vvc_rom_tables.h"

#define SCALE_BITS 15

static int ilog2(int v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* up-right diagonal order of a bw x bh grid (CL/Rom.cpp ScanGenerator, SCAN_DIAG) */
static int diag_order(int bw, int bh, int* ox, int* oy)
{
  int n = 0, d, y;
  for (d = 0; d <= bw + bh - 2; d++)
    for (y = imin(d, bh - 1); y >= imax(0, d - bw + 1); y--) { ox[n] = d - y; oy[n] = y; n++; }
  return n;
}

/* xGetICRateTS :2067-2150 (useLimitedPrefixLength == extendedPrecision == false) */
static int ic_rate_ts(int absLevel, const uint32_t par[2], const vvcb_dq_rates* r, const uint32_t sign[2], const uint32_t gt1[2], int sgn, int ricePar)
{
  int rate = (int)sign[sgn];
  if (absLevel > 1) {
    int cutoff = 2, i;
    rate += (int)gt1[1];
    rate += (int)par[(absLevel - 2) & 1];
    for (i = 0; i < 4; i++) {
      if (absLevel >= cutoff) rate += (int)r->ts_gtx[cutoff >> 1][absLevel >= cutoff + 2];
      cutoff += 2;
    }
    if (absLevel >= cutoff) {
      uint32_t symbol = (uint32_t)(absLevel - cutoff) >> 1, length;
      if (symbol < ((uint32_t)5 << ricePar)) {
        length = symbol >> ricePar;
        rate += (int)((length + 1 + ricePar) << SCALE_BITS);
      } else {
        length = (uint32_t)ricePar;
        symbol = symbol - ((uint32_t)5 << ricePar);
        while (symbol >= ((uint32_t)1 << length)) symbol -= (uint32_t)1 << (length++);
        rate += (int)((5 + length + 1 - ricePar + length) << SCALE_BITS);
      }
    }
  } else if (absLevel == 1) rate += (int)gt1[0];
  else rate = 0;
  return rate;
}

int orc_rdoq_ts(const int32_t* coeff, int w, int h, int bd, int qp, double lambda, const vvcb_dq_rates* rates, int32_t* level)
{
  static const uint8_t ricePars[32] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2 };
  const int per = qp / 6, rem = qp % 6;
  const int transformShift = 15 - bd - ((ilog2(w) + ilog2(h)) >> 1);
  const int qBits = 14 + per + transformShift;
  const int quantCoeff = kQuantScales[rem];
  /* xGetErrScaleCoeff: 2^15 * 2^(-2 * transformShift) / QStep / QStep */
  const double errorScale = ((double)(1 << SCALE_BITS) * pow(2.0, -2.0 * (double)transformShift)) / quantCoeff / quantCoeff / 1;
  const int entropyMax = (1 << 15) - 1;
  const int gw = w >> 2, gh = h >> 2, sbNum = gw * gh;
  int gx[64], gy[64], ix[16], iy[16];
  uint8_t sigGroup[64];
  int sb, absSum = 0, anySigCG = 0;
  memset(level, 0, sizeof(int32_t) * w * h);
  memset(sigGroup, 0, sizeof(sigGroup));
  diag_order(gw, gh, gx, gy);
  diag_order(4, 4, ix, iy);
  for (sb = 0; sb < sbNum; sb++) {
    const int sbPos = gy[sb] * gw + gx[sb];
    const int sigLeft = gx[sb] > 0 ? sigGroup[sbPos - 1] : 0, sigAbove = gy[sb] > 0 ? sigGroup[sbPos - gw] : 0;
    const uint32_t* bitsSigGroup = rates->ts_sig_sbb[sigLeft + sigAbove];
    int noCoeffCoded = 0, i;
    double baseCost = 0.0, sigCostSum = 0.0, codedLevelAndDist = 0.0, uncodedDist = 0.0;
    double costCoeff[16], costCoeff0[16], costSig[16];
    for (i = 0; i < 16; i++) {
      const int px = gx[sb] * 4 + ix[i], py = gy[sb] * 4 + iy[i], blk = py * w + px;
      const int64_t tmpLevel = (int64_t)abs(coeff[blk]) * quantCoeff;
      const int64_t cap = (int64_t)INT32_MAX - ((int64_t)1 << (qBits - 1));
      const int32_t levelDouble = (int32_t)(tmpLevel < cap ? tmpLevel : cap);
      const uint32_t roundAbs = (uint32_t)imin(entropyMax, (int)(((int64_t)levelDouble + ((int64_t)1 << (qBits - 1))) >> qBits));
      const uint32_t minAbs = roundAbs > 1 ? roundAbs - 1 : 1;
      const uint32_t downAbs = (uint32_t)imin(entropyMax, levelDouble >> qBits);
      const uint32_t upAbs = (uint32_t)imin(entropyMax, (int)downAbs + 1);
      uint32_t tested[3];
      int nTested = 0, right, below, pred1, predPixel, numPos, sum, ricePar, signCtx, k, isLast;
      const uint32_t *bitsSig, *bitsPar, *bitsSign, *bitsGt1;
      const int sgn = coeff[blk] < 0;
      double dErr, cost, cost0, csig, currCostSig = 0.0;
      uint32_t best = 0;
      tested[nTested++] = roundAbs;
      if (minAbs != roundAbs) tested[nTested++] = minAbs;
      right = px > 0 ? level[blk - 1] : 0;                  /* neighTS: the sample to the left ... */
      below = py > 0 ? level[blk - w] : 0;                  /* ... and the one above (names as in the reference) */
      pred1 = imax(abs(below), abs(right));
      predPixel = (int)upAbs == pred1 ? 1 : ((int)upAbs < pred1 ? (int)upAbs + 1 : (int)upAbs);    /* deriveModCoeff */
      if (upAbs != roundAbs && upAbs != minAbs && predPixel == 1) tested[nTested++] = upAbs;
      dErr = (double)levelDouble;
      cost0 = dErr * dErr * errorScale;
      costCoeff0[i] = cost0;
      level[blk] = (int32_t)tested[0];
      numPos = (right != 0) + (below != 0);
      bitsSig = rates->ts_sig[numPos];
      bitsPar = rates->ts_par[0];
      sum = abs(right) + abs(below);
      ricePar = ricePars[imin(sum, 31)];
      if ((right == 0 && below == 0) || ((int64_t)right * below < 0)) signCtx = 0;
      else if (right >= 0 && below >= 0) signCtx = 1;
      else signCtx = 2;
      bitsSign = rates->ts_sign[signCtx];
      bitsGt1 = rates->ts_lrg1[numPos];
      isLast = (i == 15 && noCoeffCoded == 0);
      /* xGetCodedLevelTSPred */
      cost = 0.0; csig = 0.0;
      if (!isLast && tested[0] < 3) {
        csig = lambda * (double)bitsSig[0];
        cost = cost0 + csig;
        if (tested[0] == 0) goto decided;
      } else cost = DBL_MAX;
      if (!isLast) currCostSig = lambda * (double)bitsSig[1];
      for (k = 0; k < nTested; k++) {
        const int absLevel = (int)tested[k];
        const double e = (double)(levelDouble - (int32_t)((uint32_t)absLevel << qBits));
        const double err = e * e * errorScale;
        const int mod = absLevel == pred1 ? 1 : (absLevel < pred1 ? absLevel + 1 : absLevel);
        double cur = err + lambda * (double)ic_rate_ts(mod, bitsPar, rates, bitsSign, bitsGt1, sgn, ricePar);
        cur += currCostSig;
        if (cur < cost) { best = (uint32_t)absLevel; cost = cur; csig = currCostSig; }
      }
decided:
      costCoeff[i] = cost; costSig[i] = csig;
      if (best > 0) noCoeffCoded++;
      level[blk] = (best != 0 && coeff[blk] < 0) ? -(int32_t)best : (int32_t)best;
      baseCost += costCoeff[i];
      sigCostSum += costSig[i];
      if (level[blk]) {
        sigGroup[sbPos] = 1;
        codedLevelAndDist += costCoeff[i] - costSig[i];
        uncodedDist += costCoeff0[i];
      }
    }
    if (!sigGroup[sbPos]) {
      baseCost += lambda * (double)bitsSigGroup[0] - sigCostSum;
    } else if (sb != sbNum - 1 || anySigCG) {
      double costZeroSB = baseCost;
      baseCost += lambda * (double)bitsSigGroup[1];
      costZeroSB += lambda * (double)bitsSigGroup[0];
      costZeroSB += uncodedDist;
      costZeroSB -= codedLevelAndDist;
      costZeroSB -= sigCostSum;
      if (costZeroSB < baseCost) {
        sigGroup[sbPos] = 0;
        for (i = 0; i < 16; i++) level[(gy[sb] * 4 + iy[i]) * w + gx[sb] * 4 + ix[i]] = 0;
      } else anySigCG = 1;
    }
  }
  for (sb = 0; sb < w * h; sb++) absSum += abs(level[sb]);
  return absSum;
}
