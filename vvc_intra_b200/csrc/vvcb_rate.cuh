// vvc_intra_b200 -- residual rate estimation kernel (sm_100a): the fractional bits CABACWriter::residual_coding( tu, COMPONENT_Y )
// adds on the reference's bit estimator (EL/CABACWriter.cpp:3773-3895, mts_coding :3897, last_sig_coeff :3960, residual_coding_subblock
// :4164, residual_codingTS :4025, residual_coding_subblockTS :4305; contexts CL/ContextModelling.h; BinProbModel_Std CL/Contexts.h:90-163;
// bypass / Golomb-Rice pricing EL/BinEncoder.cpp).  Every context-coded bin is priced from the model's current state and then adapts the
// model, so a TU is one serial chain over its bins: one thread per TU on a private copy of the context models (the host sorts the jobs
// by size).  Luma, no sign hiding (off with dependent quantisation), no BDPCM, no ISP.
#pragma once
#include "vvcb_dq.cuh"

namespace {

struct RateRom { uint32_t binFracBits[512]; };     // ProbModelTables::m_binFracBits (CL/Contexts.cpp:57): [state][bin]

struct RateParams {
  const vvcb_tu_job* jobs;
  const int* order;          // indices of the VVCB_TU_RATE jobs, largest TU first
  int n;
  const int32_t* level;      // final levels, dense per job at job.offset
  vvcb_tu_result* results;   // frac_bits is written here
  const vvcb_ctx_states* states;
  const DqRom* rom;
  const RateRom* rate;
  int depQuant;              // slice->getDepQuantEnabledFlag(): selects the state machine that picks the significance context set
};

struct RateEst {
  unsigned long long bits;
  const uint32_t* frac;
};

// TBitEstimator::encodeBin: price, then BinProbModel_Std::update (MASK_0 0x7fe0, MASK_1 0x7ffe)
__device__ __forceinline__ void rate_bin(RateEst& e, vvcb_bin_model& m, unsigned bin)
{
  const int rate0 = m.rate >> 4, rate1 = m.rate & 15;
  const unsigned s0 = m.state[0], s1 = m.state[1];
  e.bits += e.frac[2 * ((s0 + s1) >> 8) + bin];
  unsigned n0 = s0 - ((s0 >> rate0) & 0x7fe0u), n1 = s1 - ((s1 >> rate1) & 0x7ffeu);
  if (bin) { n0 += (0x7fffu >> rate0) & 0x7fe0u; n1 += (0x7fffu >> rate1) & 0x7ffeu; }
  m.state[0] = (uint16_t)n0; m.state[1] = (uint16_t)n1;
}

// BitEstimatorBase::encodeRemAbsEP with maxLog2TrDynamicRange 15
__device__ __forceinline__ void rate_rem_abs(RateEst& e, unsigned bins, unsigned rice)
{
  unsigned n;
  if (bins < (5u << rice)) n = (bins >> rice) + 1 + rice;
  else {
    unsigned prefix = 0, suffix;
    const unsigned code = (bins >> rice) - 5;
    if (code >= (1u << 12) - 1) { prefix = 12; suffix = 15; }
    else { while (code > (2u << prefix) - 2) prefix++; suffix = prefix + rice + 1; }
    n = 5 + prefix + suffix;
  }
  e.bits += (unsigned long long)n << 15;
}

constexpr int kRateThreads = 32;

__global__ void __launch_bounds__(kRateThreads) rate_kernel(RateParams P)
{
  // the thread's private context models: 1044 bytes = 261 words apart, an odd stride, so the lanes of a warp hit different banks
  __shared__ vvcb_ctx_states sModels[kRateThreads];
  const DqRom& rom = *P.rom;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P.n; t += gridDim.x * blockDim.x) {
    const int ji = P.order[t];
    const vvcb_tu_job job = P.jobs[ji];
    const int lw = job.log2w, lh = job.log2h, w = 1 << lw, h = 1 << lh;
    const DqShape shp = rom.shape[lw - 2][lh - 2];
    const DqScanPos* scan = rom.pos + shp.first;
    const uint16_t* sbbPosTab = rom.sbbPos[lw - 2][lh - 2];
    const int32_t* coeff = P.level + job.offset;
    vvcb_ctx_states& c = sModels[threadIdx.x];
    {
      const uint16_t* src = reinterpret_cast<const uint16_t*>(&P.states[job.rate_idx]);
      uint16_t* dst = reinterpret_cast<uint16_t*>(&c);
      for (int i = 0; i < (int)(sizeof(vvcb_ctx_states) / 2); i++) dst[i] = src[i];
    }
    RateEst e; e.bits = 0; e.frac = P.rate->binFracBits;
    const int mts = job.mts_idx;
    const bool tsAllowed = (job.flags & VVCB_TU_TS_ALLOWED) != 0, mtsAllowed = (job.flags & VVCB_TU_MTS_ALLOWED) != 0;

    // the reference calls residual_coding only for a coded block
    bool any = false;
    int scanPosLast = -1;
    unsigned long long sigGroups = 0;                           // by sub-block scan index
    for (int pos = 0; pos < shp.numCoeff; pos++) if (coeff[scan[pos].idx]) { scanPosLast = pos; sigGroups |= 1ull << (pos >> 4); any = true; }
    if (!any) { P.results[ji].frac_bits = 0; continue; }

    // ---- mts_coding
    if (tsAllowed) rate_bin(e, c.mts_idx[6], mts == 1);
    if (mts != 1 && mtsAllowed) {
      const unsigned symbol = mts != 0;
      rate_bin(e, c.mts_idx[0], symbol);
      if (symbol)
        for (int i = 0, ctx = 7; i < 3; i++, ctx++) {
          const unsigned s2 = mts > i + 2;
          rate_bin(e, c.mts_idx[ctx], s2);
          if (!s2) break;
        }
    }

    if (mts == 1) {
      // ---- residual_codingTS: forward over the sub-blocks, left / upper neighbours as context
      int remCtxBins = 2 * w * h;
      unsigned long long coded = 0;                             // m_sigCoeffGroupFlag, by raster position of the sub-block
      const int lastRaster = shp.numSbb - 1;
      for (int sb = 0; sb < shp.numSbb; sb++) {
        const int sbPos = sbbPosTab[sb], sy = sbPos / shp.widthInSbb, sx = sbPos - sy * shp.widthInSbb;
        const bool sig = (sigGroups >> sb) & 1;
        if (sig) coded |= 1ull << sbPos;
        const int sigLeft = sx > 0 ? (int)((coded >> (sbPos - 1)) & 1) : 0, sigAbove = sy > 0 ? (int)((coded >> (sbPos - shp.widthInSbb)) & 1) : 0;
        if (sb != shp.numSbb - 1 || (coded & ~(1ull << lastRaster)) != 0) {
          rate_bin(e, c.ts_sig_sbb[sigLeft + sigAbove], sig);
          if (!sig) continue;
        }
        int numNonZero = 0;
        for (int i = 0; i < 16; i++) {
          const DqScanPos sp = scan[sb * 16 + i];
          const int v = coeff[sp.idx];
          const int left = sp.x > 0 ? coeff[sp.idx - 1] : 0, above = sp.y > 0 ? coeff[sp.idx - w] : 0;
          const int numPos = (left != 0) + (above != 0);
          if (numNonZero || i != 15) { if (--remCtxBins >= 0) rate_bin(e, c.ts_sig[numPos], v != 0); else e.bits += 1ull << 15; }
          if (v) {
            int signCtx;
            if ((left == 0 && above == 0) || ((long long)left * above < 0)) signCtx = 0;
            else if (left >= 0 && above >= 0) signCtx = 1;
            else signCtx = 2;
            if (--remCtxBins >= 0) rate_bin(e, c.ts_sign[signCtx], v < 0); else e.bits += 1ull << 15;
            numNonZero++;
            const int pred1 = vmax(vabs(left), vabs(above)), a = vabs(v);
            int rem = (a == pred1 ? 1 : (a < pred1 ? a + 1 : a)) - 1;
            if (--remCtxBins >= 0) rate_bin(e, c.ts_lrg1[numPos], rem != 0); else e.bits += 1ull << 15;
            if (rem) {
              rem -= 1;
              if (--remCtxBins >= 0) rate_bin(e, c.ts_par[0], rem & 1); else e.bits += 1ull << 15;
            }
          }
        }
        for (int i = 0; i < 16; i++) {
          const DqScanPos sp = scan[sb * 16 + i];
          const int left = sp.x > 0 ? coeff[sp.idx - 1] : 0, above = sp.y > 0 ? coeff[sp.idx - w] : 0;
          const int pred1 = vmax(vabs(left), vabs(above)), a = vabs(coeff[sp.idx]);
          const int mod = a == pred1 ? 1 : (a < pred1 ? a + 1 : a);
          for (int k = 0, cutoff = 2; k < 4; k++, cutoff += 2)
            if (mod >= cutoff) { if (--remCtxBins >= 0) rate_bin(e, c.ts_gtx[cutoff >> 1], mod >= cutoff + 2); else e.bits += 1ull << 15; }
        }
        for (int i = 0; i < 16; i++) {
          const DqScanPos sp = scan[sb * 16 + i];
          const int left = sp.x > 0 ? coeff[sp.idx - 1] : 0, above = sp.y > 0 ? coeff[sp.idx - w] : 0;
          const int pred1 = vmax(vabs(left), vabs(above)), a = vabs(coeff[sp.idx]);
          const int mod = a == pred1 ? 1 : (a < pred1 ? a + 1 : a);
          if (mod >= 10) rate_rem_abs(e, (unsigned)(mod - 10) >> 1, rom.tsRicePars[vmin(vabs(left) + vabs(above), 31)]);
        }
      }
      P.results[ji].frac_bits = e.bits;
      continue;
    }

    // ---- last_sig_coeff
    {
      const DqScanPos lp = scan[scanPosLast];
      const int gX = rom.groupIdx[lp.x], gY = rom.groupIdx[lp.y];
      int maxX = rom.groupIdx[vmin(32, w) - 1], maxY = rom.groupIdx[vmin(32, h) - 1];
      if (mts > 1) { if (w == 32) maxX = rom.groupIdx[15]; if (h == 32) maxY = rom.groupIdx[15]; }
      const int offX = lw == 2 ? 0 : lw == 3 ? 3 : lw == 4 ? 6 : lw == 5 ? 10 : 15, offY = lh == 2 ? 0 : lh == 3 ? 3 : lh == 4 ? 6 : lh == 5 ? 10 : 15;
      const int shX = (lw + 1) >> 2, shY = (lh + 1) >> 2;
      for (int k = 0; k < gX; k++) rate_bin(e, c.last_x[offX + (k >> shX)], 1);
      if (gX < maxX) rate_bin(e, c.last_x[offX + (gX >> shX)], 0);
      for (int k = 0; k < gY; k++) rate_bin(e, c.last_y[offY + (k >> shY)], 1);
      if (gY < maxY) rate_bin(e, c.last_y[offY + (gY >> shY)], 0);
      if (gX > 3) e.bits += (unsigned long long)((gX - 2) >> 1) << 15;
      if (gY > 3) e.bits += (unsigned long long)((gY - 2) >> 1) << 15;
    }

    // ---- sub-blocks from the last significant one down to the DC one
    int regBins;
    {
      int tbW = w, tbH = h;
      if (mts > 1) { tbW = w == 32 ? 16 : w; tbH = h == 32 ? 16 : h; }
      regBins = (vmin(32, tbW) * vmin(32, tbH) * 28) >> 4;       // TU::getTbAreaAfterCoefZeroOut * 28 >> 4
    }
    const int stateTab = P.depQuant ? 32040 : 0;
    int state = 0;
    unsigned long long coded = 0;                                // by raster position
    for (int sb = scanPosLast >> 4; sb >= 0; sb--) {
      const int sbPos = sbbPosTab[sb], sy = sbPos / shp.widthInSbb, sx = sbPos - sy * shp.widthInSbb, minSub = sb * 16;
      const bool sig = (sigGroups >> sb) & 1, isLast = (scanPosLast >> 4) == sb;
      if (sig) coded |= 1ull << sbPos;
      if (mts > 1 && ((h == 32 && sy >= 4) || (w == 32 && sx >= 4))) continue;
      if (!isLast && sb != 0) {
        const int sigRight = sx + 1 < shp.widthInSbb ? (int)((coded >> (sbPos + 1)) & 1) : 0;
        const int sigLower = sy + 1 < shp.heightInSbb ? (int)((coded >> (sbPos + shp.widthInSbb)) & 1) : 0;
        rate_bin(e, c.sig_sbb[sigRight | sigLower], sig);
        if (!sig) continue;
      }
      const int firstSigPos = isLast ? scanPosLast : minSub + 15;
      int next = firstSigPos, numNonZero = 0, signBins = 0;
      const int inferSigPos = next != scanPosLast ? (sb != 0 ? minSub : -1) : next;
      for (; next >= minSub && regBins >= 4; next--) {
        const DqScanPos sp = scan[next];
        const int v = coeff[sp.idx];
        const int diag = sp.x + sp.y;
        int sumAbs1 = 0, numPos = 0;
        {
          const int32_t* p = coeff + sp.idx;
#define VVCB_T(q) { const int a = vabs(q); sumAbs1 += vmin(4 + (a & 1), a); numPos += a != 0; }
          if (sp.x < w - 1) { VVCB_T(p[1]); if (sp.x < w - 2) VVCB_T(p[2]); if (sp.y < h - 1) VVCB_T(p[w + 1]); }
          if (sp.y < h - 1) { VVCB_T(p[w]); if (sp.y < h - 2) VVCB_T(p[2 * w]); }
#undef VVCB_T
        }
        if (numNonZero || next != inferSigPos) {
          rate_bin(e, c.sig[vmax(0, state - 1)][vmin((sumAbs1 + 1) >> 1, 3) + (diag < 2 ? 4 : 0) + (diag < 5 ? 4 : 0)], v != 0);
          regBins--;
        }
        if (v) {
          // ctxOffsetAbs(): template of this position, except for the TU's last position, whose significance context is never derived
          const int ctxOff = next != scanPosLast ? vmin(sumAbs1 - numPos, 4) + 1 + (diag == 0 ? 15 : diag < 3 ? 10 : diag < 10 ? 5 : 0) : 0;
          int rem = vabs(v) - 1;
          numNonZero++; signBins++;
          rate_bin(e, c.gt1[ctxOff], rem != 0);
          regBins--;
          if (rem) {
            rem -= 1;
            rate_bin(e, c.par[ctxOff], rem & 1);
            rem >>= 1;
            regBins--;
            rate_bin(e, c.gt2[ctxOff], rem != 0);
            regBins--;
          }
        }
        state = (stateTab >> ((state << 2) + ((v & 1) << 1))) & 3;
      }
      const int firstPosMode2 = next;
      for (int i = firstSigPos; i >= minSub; i--) {
        const DqScanPos sp = scan[i];
        const int a = vabs(coeff[sp.idx]);
        if (i > firstPosMode2 && a < 4) continue;                // 2nd pass codes a remainder only from level 4 on
        int sumAbs = 0;
        {
          const int32_t* p = coeff + sp.idx;
          if (sp.x < w - 1) { sumAbs += vabs(p[1]); if (sp.x < w - 2) sumAbs += vabs(p[2]); if (sp.y < h - 1) sumAbs += vabs(p[w + 1]); }
          if (sp.y < h - 1) { sumAbs += vabs(p[w]); if (sp.y < h - 2) sumAbs += vabs(p[2 * w]); }
        }
        if (i > firstPosMode2) rate_rem_abs(e, (unsigned)(a - 4) >> 1, rom.goRicePars[vmax(vmin(sumAbs - 20, 31), 0)]);
        else {                                                   // bypass-coded coefficient
          const int sumAll = vmin(sumAbs, 31);
          const int pos0 = rom.goRiceZero[vmax(0, state - 1)][sumAll];
          rate_rem_abs(e, (unsigned)(a == 0 ? pos0 : (a <= pos0 ? a - 1 : a)), rom.goRicePars[sumAll]);
          state = (stateTab >> ((state << 2) + ((a & 1) << 1))) & 3;
          if (a) signBins++;
        }
      }
      e.bits += (unsigned long long)signBins << 15;
    }
    P.results[ji].frac_bits = e.bits;
  }
}

}  // namespace
