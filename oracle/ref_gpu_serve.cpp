// TEST / INTEGRATION INFRASTRUCTURE -- the UNMODIFIED reference encoder with the luma intra cost evaluation SERVED by
// libvvc_intra_b200.so (oracle/_ref/EncoderAppServe, built by oracle/Makefile.ref).
//
// Where oracle/ref_gpu_shim.cpp runs the reference's own arithmetic first and compares (a shadow), this shim REPLACES it: inside
// IntraSearch::estIntraPredLumaQT (EL/IntraSearch.cpp:289) the reference no longer fetches reference samples, predicts, measures
// SAD / SATD, transforms, quantises, reconstructs, measures the SSE or prices the residual of a luma TU that covers its CU -- the
// wrappers below return the engine's numbers and the reference's control flow (candidate lists, early outs, mode decision, CABAC
// state, split search) consumes them.  The reference sources are not edited; the seams are the cross-object calls `ld --wrap` can
// intercept (S1-S4 of SURVEY.md 8b):
//
//   IntraSearch::estIntraPredLumaQT          per call: ONE vvcb_cu_eval round trip (two on the first pass of a CU: lists, then TUs)
//     that pushes the reconstructed neighbourhood, runs the rough mode decision (every SAD / SATD of the visit) and codes every
//     (mode, transform, LFNST) candidate the full-RD loop of this pass can reach -- a prefetch; what the loop then really asks
//     for is looked up.  A lookup is valid only if the neighbourhood, QP, lambda, cbf price and the CABAC context snapshot it was
//     computed with are the ones of the asking call (hashes / values are compared), otherwise the candidate is fetched on demand:
//     speculation decides the hit rate, never the result.
//   IntraPrediction::initIntraPatternChType / initIntraMip     skipped (the engine builds the reference lines itself)
//   IntraPrediction::predIntraAng / predIntraMip               RMD: nothing to compute; full RD: the engine's samples
//   RdCost::setDistParam                                       RMD: SAD / SATD of the slot under test from vvcb_rmd_detail
//   TrQuant::transformNxN (both), invTransformNxN              pre-selection sums, levels + absSum, residual = reco - pred
//   RdCost::getDistPart (DF_SSE)                               vvcb_tu_result::sse
//   CABACWriter::residual_coding on the estimator              vvcb_tu_result::frac_bits
//
// ISP sub-partitions, BDPCM, chroma and everything outside estIntraPredLumaQT run the reference's own code.  With VVCB_BROKER set
// the context is a broker client (include/vvc_intra_b200_broker.h): many such encoder processes share one GPU context.
// tests/test_serve_shim.py (CPU, oracle-backed engine) and tests/test_gpu_parity.py (-m gpu) require byte-identical bitstreams.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
#include <map>
#include <list>
#include <set>
#include <array>
#include <algorithm>
#include <functional>
#include <memory>
#include <sstream>
#include <iostream>
#include <fstream>
#include <mutex>
#include <cmath>
#include <limits>
#include <deque>
#include <bitset>
#include <unordered_map>
#include <atomic>
#include <chrono>
#include <iomanip>
#include <cassert>
#include <numeric>
#include <stack>
#include <stdexcept>
#include <utility>
#include <type_traits>
#include <exception>
#include <iterator>
#include <tuple>
#include <cstdarg>
#include <cstddef>
#include <climits>

#define private public
#define protected public
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/Picture.h"
#include "CommonLib/IntraPrediction.h"
#include "CommonLib/RdCost.h"
#include "CommonLib/TrQuant.h"
#include "CommonLib/Quant.h"
#include "CommonLib/Contexts.h"
#include "CommonLib/ContextModelling.h"
#include "CommonLib/UnitPartitioner.h"
#include "EncoderLib/CABACWriter.h"
#include "EncoderLib/BinEncoder.h"
#include "EncoderLib/IntraSearch.h"
#include "EncoderLib/EncCfg.h"
#undef private
#undef protected

#include "../include/vvc_intra_b200.h"

namespace {

// ---- one candidate TU of the current CU, as the engine returned it --------------------------------------------------------------
struct TuEntry {
  vvcb_tu_job    job;
  vvcb_tu_result res;
  std::vector<int32_t> level;               // own storage (candidates fetched one by one); the first-pass prefetch leaves its blocks in
  std::vector<int16_t> reco;                //   the shim's reusable arrays and only sets the pointers
  const int32_t* lv = nullptr;              // the levels / the reconstruction of this candidate, w * h each
  const int16_t* rc = nullptr;
  uint64_t rateHash = 0, stateHash = 0;     // context snapshot the levels / the bits were computed with
  int      slot = 0;
};

struct CuCache {
  bool     valid = false;
  int      x = 0, y = 0, w = 0, h = 0;
  uint64_t nbhHash = 0;
  bool     pushed = false;                  // neighbourhood rectangles are on the device
  bool     rmdValid = false;
  vvcb_rmd_visit  visit;
  vvcb_rmd_result res;
  vvcb_rmd_detail det;
  std::map<uint32_t, TuEntry> tus;          // key: slot | lfnst << 8 | mtsIdx << 12
  std::map<int, std::vector<int16_t>> pred; // per slot
  std::vector<vvcb_rect> rects;
  std::vector<int16_t>   rectSamples;
};

vvcb_ctx* g_gpu = nullptr;
int       g_poc = -1 << 30;
int       g_ctu = 128;
bool      g_enabled = true;                 // false: configuration the engine does not cover -> everything runs the reference's code
CuCache   g_cu;
bool      g_inEst = false;                  // inside a served estIntraPredLumaQT
bool      g_inRmd = false;                  // between initIntraPatternChType( forceRefFilterFlag ) and the first full-RD prediction
int       g_curSlot = -1;                   // slot of the last skipped RMD prediction
IntraSearch* g_is = nullptr;

// VVCB_SHIM_OFF=1 (every call runs the reference's own code): the wrappers only count the rough-mode-decision evaluations of the plain
// encoder -- one prediction + SAD + SATD each -- for bench.py's reference arm (BASELINE.json's second metric, SATD block-mode evals / s)
bool g_counting = false, g_countRmd = false;
int  g_cntX = 0, g_cntY = 0, g_cntW = 0, g_cntH = 0;
long g_satdEvals = 0, g_tuQuantReal = 0;
bool countArea(const CompArea& a) { return g_counting && a.compID == COMPONENT_Y && a.x == g_cntX && a.y == g_cntY && (int)a.width == g_cntW && (int)a.height == g_cntH; }

struct PendingTu { const TransformUnit* tu = nullptr; TuEntry* e = nullptr; const Pel* recoBuf = nullptr; } g_pend;

struct Tm { long long ns = 0; };
struct Scope { Tm& t; std::chrono::steady_clock::time_point t0; Scope(Tm& x) : t(x), t0(std::chrono::steady_clock::now()) {} ~Scope() { t.ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); } };
Tm g_tEst, g_tPre, g_tWrap;
// the served wrappers run a few million times per CTU: their own time is only clocked on request (two clock reads per call were 8 % of a walker's CPU time)
const bool g_wrapTimed = getenv("VVCB_SHIM_TIMES") != nullptr;
struct WrapScope { long long t0; WrapScope() : t0(g_wrapTimed ? std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0) {}
                   ~WrapScope() { if (g_wrapTimed) g_tWrap.ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0; } };
Tm g_prof[8][3];      // VVCB_SHIM_PROFILE=1: time inside the REAL functions by [family][luma whole-CU, luma ISP, chroma]
const bool g_profile = getenv("VVCB_SHIM_PROFILE") != nullptr;
const char* kFamily[8] = { "ref_fetch", "predict", "dist_param", "tr_presel", "tr_quant", "inv_tr", "sse", "resid_bits" };
inline int catOf(ComponentID c, const CodingUnit& cu) { return c != COMPONENT_Y ? 2 : (cu.ispMode ? 1 : 0); }
struct Stats {
  long estCalls = 0, visits = 0, rmdRoundTrips = 0, tuRoundTrips = 0, demandRoundTrips = 0, jobsPrefetched = 0, jobsDemand = 0;
  long predSkipped = 0, predServed = 0, distServed = 0, preselServed = 0, quantServed = 0, quantDq = 0, quantTs = 0, quantLfnst = 0;
  long invServed = 0, sseServed = 0, bitsServed = 0, bitsReal = 0, refFetchSkipped = 0, staleRate = 0, cuReuse = 0;
  long long engineNs = 0;
} g_st;

void die(const char* what, const char* detail = "")
{
  fprintf(stderr, "vvcb serve shim: %s %s\n", what, detail);
  fflush(stderr);
  abort();
}
void gpuCheck(int rc, const char* what) { if (rc != VVCB_OK) die(what, vvcb_last_error(g_gpu)); }

void report()
{
  if (const char* p = getenv("VVCB_SHIM_REPORT"))
    if (FILE* f = fopen(p, "w")) {
      fprintf(f, "{\"mode\": \"serve\", \"enabled\": %d, \"est_calls\": %ld, \"visits\": %ld, \"cu_reuse\": %ld, \"rmd_round_trips\": %ld, \"tu_round_trips\": %ld, \"demand_round_trips\": %ld, "
                 "\"jobs_prefetched\": %ld, \"jobs_on_demand\": %ld, \"ref_fetch_skipped\": %ld, \"predictions_skipped\": %ld, \"predictions_served\": %ld, \"distortions_served\": %ld, "
                 "\"tu_preselections\": %ld, \"tu_quantised\": %ld, \"tu_dep_quant\": %ld, \"tu_rdoq_ts\": %ld, \"tu_lfnst\": %ld, \"tu_reconstructions\": %ld, \"tu_sse\": %ld, "
                 "\"tu_residual_bits\": %ld, \"tu_residual_bits_reference\": %ld, \"stale_context\": %ld, \"engine_wait_s\": %.3f, \"est_total_s\": %.3f, \"est_prepare_s\": %.3f, \"wrappers_s\": %.3f",
              (int)g_enabled, g_st.estCalls, g_st.visits, g_st.cuReuse, g_st.rmdRoundTrips, g_st.tuRoundTrips, g_st.demandRoundTrips, g_st.jobsPrefetched, g_st.jobsDemand,
              g_st.refFetchSkipped, g_st.predSkipped, g_st.predServed, g_st.distServed, g_st.preselServed, g_st.quantServed, g_st.quantDq, g_st.quantTs, g_st.quantLfnst,
              g_st.invServed, g_st.sseServed, g_st.bitsServed, g_st.bitsReal, g_st.staleRate, g_st.engineNs * 1e-9, g_tEst.ns * 1e-9, g_tPre.ns * 1e-9, g_tWrap.ns * 1e-9);
      fprintf(f, ", \"reference_satd_evals\": %ld, \"reference_tu_quantisations\": %ld", g_satdEvals, g_tuQuantReal);
      if (g_gpu && getenv("VVCB_SHIM_KERNEL_TIMES")) {
        float t[4] = { 0, 0, 0, 0 }, r[3] = { 0, 0, 0 }; int calls = 0, launches = 0;
        vvcb_tu_kernel_times(g_gpu, t, &calls);
        vvcb_kernel_times(g_gpu, r, &launches);
        fprintf(f, ", \"tu_kernel_ms\": [%.1f, %.1f, %.1f, %.1f], \"tu_calls\": %d, \"rmd_kernel_ms\": [%.1f, %.1f, %.1f], \"rmd_calls\": %d", t[0], t[1], t[2], t[3], calls, r[0], r[1], r[2], launches);
      }
      if (g_profile) for (int k = 0; k < 8; k++) fprintf(f, ", \"real_%s_s\": [%.3f, %.3f, %.3f]", kFamily[k], g_prof[k][0].ns * 1e-9, g_prof[k][1].ns * 1e-9, g_prof[k][2].ns * 1e-9);
      fprintf(f, "}\n");
      fclose(f);
    }
  if (g_gpu) vvcb_destroy(g_gpu);
}

uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull)
{
  const uint8_t* b = static_cast<const uint8_t*>(p);
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
// word-wise variant for the larger blobs (context snapshots, neighbourhood samples)
uint64_t mix(const void* p, size_t n, uint64_t h = 0x9e3779b97f4a7c15ull)
{
  const uint8_t* b = static_cast<const uint8_t*>(p);
  size_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t v; memcpy(&v, b + i, 8); h = (h ^ v) * 0xff51afd7ed558ccdull; h ^= h >> 32; }
  for (; i < n; i++) { h = (h ^ b[i]) * 0xff51afd7ed558ccdull; h ^= h >> 29; }
  return h;
}

void ensureFrame(const CodingStructure& cs)
{
  const SPS& sps = *cs.sps;
  if (!g_gpu) {
    g_ctu = sps.getMaxCUWidth();
    if (vvcb_create(&g_gpu, 0, sps.getBitDepth(CHANNEL_TYPE_LUMA), g_ctu) != VVCB_OK) die("vvcb_create:", vvcb_last_error(nullptr));
    gpuCheck(vvcb_set_option(g_gpu, VVCB_OPT_DEP_QUANT, cs.slice->getDepQuantEnabledFlag() ? 1 : 0), "vvcb_set_option:");
    if (getenv("VVCB_SHIM_KERNEL_TIMES")) vvcb_kernel_timing(g_gpu, 1);          // profiling aid: CUDA events around the kernel stages of every call
    atexit(report);
  }
  if (cs.slice->getPOC() != g_poc) {
    g_poc = cs.slice->getPOC();
    const CPelBuf org = cs.picture->getOrigBuf(COMPONENT_Y);          // the LMCS-mapped original, constant during the CTU loop (EL/EncGOP.cpp:1692)
    gpuCheck(vvcb_frame_begin(g_gpu, org.buf, org.stride, org.width, org.height), "vvcb_frame_begin:");
    g_cu.valid = false;
  }
}

bool unitAvail(const CodingStructure& cs, const CodingUnit& cu, const Position& p)
{
  return cs.isDecomp(p, CH_L) && cs.getCURestricted(p, cu, CH_L) != nullptr;
}

void fillRates(vvcb_dq_rates& out, const Ctx& ctx)
{
  const FracBitsAccess& fb = ctx.getFracBitsAcess();
  uint32_t* p = reinterpret_cast<uint32_t*>(&out);
  auto put = [&](const CtxSet& set, int num) { for (int i = 0; i < num; i++) { const BinFracBits b = fb.getFracBitsArray(set(i)); *p++ = b.intBits[0]; *p++ = b.intBits[1]; } };
  put(Ctx::SigCoeffGroup[CHANNEL_TYPE_LUMA], 2);
  for (int st = 0; st < 3; st++) put(Ctx::SigFlag[CHANNEL_TYPE_LUMA + 2 * st], 12);
  put(Ctx::ParFlag[CHANNEL_TYPE_LUMA], 21); put(Ctx::GtxFlag[2 + CHANNEL_TYPE_LUMA], 21); put(Ctx::GtxFlag[CHANNEL_TYPE_LUMA], 21);
  put(Ctx::LastX[CHANNEL_TYPE_LUMA], 20); put(Ctx::LastY[CHANNEL_TYPE_LUMA], 20);
  put(Ctx::TsSigCoeffGroup, 3); put(Ctx::TsSigFlag, 3); put(Ctx::TsParFlag, 1); put(Ctx::TsGtxFlag, 5); put(Ctx::TsLrg1Flag, 4); put(Ctx::TsResidualSign, 6);
  if ((char*)p != (char*)&out + sizeof(out)) die("vvcb_dq_rates layout");
}

void fillStates(vvcb_ctx_states& out, const Ctx& ctx)
{
  vvcb_bin_model* p = reinterpret_cast<vvcb_bin_model*>(&out);
  auto put = [&](const CtxSet& set, int num) { for (int i = 0; i < num; i++) { const BinProbModel_Std& m = ctx.m_CtxStore_Std[set(i)]; p->state[0] = m.m_state[0]; p->state[1] = m.m_state[1]; p->rate = m.m_rate; p->pad = 0; p++; } };
  put(Ctx::MTSIndex, 11);
  put(Ctx::SigCoeffGroup[CHANNEL_TYPE_LUMA], 2);
  for (int k = 0; k < 3; k++) put(Ctx::SigFlag[CHANNEL_TYPE_LUMA + 2 * k], 12);
  put(Ctx::ParFlag[CHANNEL_TYPE_LUMA], 21); put(Ctx::GtxFlag[2 + CHANNEL_TYPE_LUMA], 21); put(Ctx::GtxFlag[CHANNEL_TYPE_LUMA], 21);
  put(Ctx::LastX[CHANNEL_TYPE_LUMA], 20); put(Ctx::LastY[CHANNEL_TYPE_LUMA], 20);
  put(Ctx::TsSigCoeffGroup, 3); put(Ctx::TsSigFlag, 3); put(Ctx::TsParFlag, 1); put(Ctx::TsGtxFlag, 5); put(Ctx::TsLrg1Flag, 4); put(Ctx::TsResidualSign, 6);
  if ((char*)p != (char*)&out + sizeof(out)) die("vvcb_ctx_states layout");
}

int cbfDeltaBits(const Ctx& ctx)
{
  const BinFracBits cbf = ctx.getFracBitsAcess().getFracBitsArray(Ctx::QtCbf[COMPONENT_Y](DeriveCtx::CtxQtCbf(COMPONENT_Y, false)));
  return int32_t(cbf.intBits[1]) - int32_t(cbf.intBits[0]);            // RateEstimator::xSetLastCoeffOffset, CL/DepQuant.cpp:531-540 (no ISP)
}

int slotOfMode(const CuCache& cu, bool mip, int mrl, int mode)
{
  if (mip) return VVCB_SLOT_MIP + mode;
  if (mrl == 0) return mode;
  for (int i = 1; i < 6; i++)
    if (cu.visit.mpm[i] == mode) return (mrl == 1 ? VVCB_SLOT_MRL1 : VVCB_SLOT_MRL3) + i - 1;
  return -1;
}
int slotOf(const PredictionUnit& pu, bool mip) { return slotOfMode(g_cu, mip, pu.multiRefIdx, pu.intraDir[0]); }

bool slotEvaluated(const CuCache& cu, int slot)
{
  if (slot < 0 || slot >= VVCB_NUM_SLOTS) return false;
  if (slot < VVCB_SLOT_MRL1) return true;
  if (slot < VVCB_SLOT_MIP) return (cu.y & (g_ctu - 1)) != 0;
  const bool mipOn = !(cu.visit.flags & VVCB_VISIT_NO_MIP) && mipModesAvailable(Size(cu.w, cu.h));
  return mipOn && slot - VVCB_SLOT_MIP < getNumModesMip(Size(cu.w, cu.h));
}

uint32_t tuKey(int slot, int lfnst, int mts) { return (uint32_t)slot | (uint32_t)lfnst << 8 | (uint32_t)mts << 12; }

struct TuSpec { int slot, lfnst, mts, intraMode; };

// ---- the engine request of the current CU: rectangles (once), lists (once), the TU candidates not yet cached ------------------------
struct Snapshot { vvcb_dq_rates rates; vvcb_ctx_states states; uint64_t rateHash, stateHash; int cbfDelta; double lambda; int per[2], rem[2]; };

void takeSnapshot(Snapshot& s, const CodingUnit& cu, const Ctx& ctx)
{
  fillRates(s.rates, ctx);
  fillStates(s.states, ctx);
  s.rateHash = mix(&s.rates, sizeof(s.rates));
  s.stateHash = mix(&s.states, sizeof(s.states));
  s.cbfDelta = cbfDeltaBits(ctx);
  s.lambda = g_is->m_pcTrQuant->m_quant->m_lambdas[COMPONENT_Y];                 // TrQuant::selectLambda( COMPONENT_Y ), EL/IntraSearch.cpp:2876
  const SPS& sps = *cu.cs->sps;
  const QpParam qp(cu.qp, COMPONENT_Y, sps.getQpBDOffset(CHANNEL_TYPE_LUMA), sps.getMinQpPrimeTsMinus4(CHANNEL_TYPE_LUMA), 0, cu.chromaFormat, 0, &sps);   // QpParam( tu, COMPONENT_Y ), CL/Quant.cpp:139
  for (int t = 0; t < 2; t++) { s.per[t] = qp.per(t != 0); s.rem[t] = qp.rem(t != 0); }
}

// The contexts a luma TU reads (quantiser prices, residual_coding states, the cbf price) as they were at the entry of the current
// estIntraPredLumaQT call: the reference restores the estimator to that state before every candidate, so nearly every lookup sees
// them unchanged -- one short compare instead of a fresh snapshot.
std::vector<unsigned>          g_resCtxIds;
std::vector<BinProbModel_Std>  g_entryModels;
Snapshot                       g_entrySnap;

void buildResidualCtxIds()
{
  if (!g_resCtxIds.empty()) return;
  auto put = [&](const CtxSet& set, int num) { for (int i = 0; i < num; i++) g_resCtxIds.push_back(set(i)); };
  put(Ctx::MTSIndex, 11);
  put(Ctx::SigCoeffGroup[CHANNEL_TYPE_LUMA], 2);
  for (int k = 0; k < 3; k++) put(Ctx::SigFlag[CHANNEL_TYPE_LUMA + 2 * k], 12);
  put(Ctx::ParFlag[CHANNEL_TYPE_LUMA], 21); put(Ctx::GtxFlag[2 + CHANNEL_TYPE_LUMA], 21); put(Ctx::GtxFlag[CHANNEL_TYPE_LUMA], 21);
  put(Ctx::LastX[CHANNEL_TYPE_LUMA], 20); put(Ctx::LastY[CHANNEL_TYPE_LUMA], 20);
  put(Ctx::TsSigCoeffGroup, 3); put(Ctx::TsSigFlag, 3); put(Ctx::TsParFlag, 1); put(Ctx::TsGtxFlag, 5); put(Ctx::TsLrg1Flag, 4); put(Ctx::TsResidualSign, 6);
  g_resCtxIds.push_back(Ctx::QtCbf[COMPONENT_Y](DeriveCtx::CtxQtCbf(COMPONENT_Y, false)));
}

void rememberEntryCtx(const CodingUnit& cu, const Ctx& ctx)
{
  buildResidualCtxIds();
  g_entryModels.resize(g_resCtxIds.size());
  for (size_t i = 0; i < g_resCtxIds.size(); i++) g_entryModels[i] = ctx.m_CtxStore_Std[g_resCtxIds[i]];
  takeSnapshot(g_entrySnap, cu, ctx);
}

bool sameAsEntryCtx(const Ctx& ctx)
{
  for (size_t i = 0; i < g_resCtxIds.size(); i++) {
    const BinProbModel_Std& a = ctx.m_CtxStore_Std[g_resCtxIds[i]];
    const BinProbModel_Std& b = g_entryModels[i];
    if (a.m_state[0] != b.m_state[0] || a.m_state[1] != b.m_state[1] || a.m_rate != b.m_rate) return false;
  }
  return true;
}

void makeJob(vvcb_tu_job& j, const CuCache& cu, const TuSpec& t, const Snapshot& s, bool tsAllowed, bool mtsAllowed, int index)
{
  memset(&j, 0, sizeof(j));
  const bool ts = t.mts == MTS_SKIP;
  j.x = cu.x; j.y = cu.y; j.log2w = floorLog2(cu.w); j.log2h = floorLog2(cu.h);
  j.mts_idx = t.mts;
  j.flags = VVCB_TU_QUANT | VVCB_TU_RATE | (ts ? VVCB_TU_RDOQ_TS : VVCB_TU_DEPQUANT) | (tsAllowed ? VVCB_TU_TS_ALLOWED : 0) | (mtsAllowed ? VVCB_TU_MTS_ALLOWED : 0);
  j.qp_per = s.per[ts]; j.qp_rem = s.rem[ts];
  j.offset = (uint32_t)index * cu.w * cu.h;
  j.rate_idx = 0;
  j.lfnst_idx = ts ? 0 : t.lfnst;
  j.intra_mode = t.intraMode;
  j.cbf_delta_bits = s.cbfDelta;
  j.lambda = s.lambda;
}

// fetch `specs` (those missing from the cache) in one round trip, together with the lists if they are still missing
void fetch(CuCache& cu, bool wantRmd, const std::vector<TuSpec>& specs, const Snapshot& s, bool tsAllowed, bool mtsAllowed, bool demand)
{
  std::vector<TuSpec> need;
  for (const TuSpec& t : specs) {
    auto it = cu.tus.find(tuKey(t.slot, t.lfnst, t.mts));
    if (it != cu.tus.end() && it->second.rateHash == s.rateHash && it->second.stateHash == s.stateHash && it->second.job.lambda == s.lambda &&
        it->second.job.cbf_delta_bits == s.cbfDelta && it->second.job.qp_per == s.per[t.mts == MTS_SKIP] && it->second.job.qp_rem == s.rem[t.mts == MTS_SKIP]) continue;
    bool dup = false;
    for (const TuSpec& q : need) dup = dup || (q.slot == t.slot && q.lfnst == t.lfnst && q.mts == t.mts);
    if (!dup) need.push_back(t);
  }
  wantRmd = wantRmd && !cu.rmdValid;
  if (!wantRmd && need.empty() && cu.pushed) return;
  const int n = (int)need.size(), bs = cu.w * cu.h;
  std::vector<vvcb_tu_job> jobs(n);
  std::vector<uint8_t> slots(n);
  for (int i = 0; i < n; i++) { makeJob(jobs[i], cu, need[i], s, tsAllowed, mtsAllowed, i); slots[i] = (uint8_t)need[i].slot; }
  std::vector<int32_t> level((size_t)n * bs);
  std::vector<int16_t> reco((size_t)n * bs), pred((size_t)n * bs);
  std::vector<vvcb_tu_result> res(n);
  vvcb_cu_request q;
  memset(&q, 0, sizeof(q));
  if (!cu.pushed) { q.rects = cu.rects.data(); q.n_rects = (int)cu.rects.size(); q.rect_samples = cu.rectSamples.data(); q.n_rect_samples = cu.rectSamples.size(); }
  q.visit = &cu.visit;
  q.want_rmd = wantRmd;
  q.result = &cu.res; q.detail = &cu.det;
  if (n) { q.jobs = jobs.data(); q.slots = slots.data(); q.n_jobs = n; q.rates = &s.rates; q.states = &s.states; q.level = level.data(); q.reco = reco.data(); q.pred = pred.data(); q.tu_results = res.data(); }
  const auto t0 = std::chrono::steady_clock::now();
  gpuCheck(vvcb_cu_eval(g_gpu, &q, 1), "vvcb_cu_eval:");
  g_st.engineNs += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  cu.pushed = true;
  if (wantRmd) { cu.rmdValid = true; g_st.rmdRoundTrips++; g_st.visits++; }
  if (n) { (demand ? g_st.demandRoundTrips : g_st.tuRoundTrips)++; (demand ? g_st.jobsDemand : g_st.jobsPrefetched) += n; }
  for (int i = 0; i < n; i++) {
    TuEntry& e = cu.tus[tuKey(need[i].slot, need[i].lfnst, need[i].mts)];
    e.job = jobs[i]; e.res = res[i]; e.slot = need[i].slot;
    e.level.assign(level.begin() + (size_t)i * bs, level.begin() + (size_t)(i + 1) * bs);
    e.reco.assign(reco.begin() + (size_t)i * bs, reco.begin() + (size_t)(i + 1) * bs);
    e.lv = e.level.data(); e.rc = e.reco.data();
    e.rateHash = s.rateHash; e.stateHash = s.stateHash;
    std::vector<int16_t>& p = cu.pred[need[i].slot];
    if (p.empty()) p.assign(pred.begin() + (size_t)i * bs, pred.begin() + (size_t)(i + 1) * bs);
  }
}

// First pass of a CU in ONE round trip: the rectangles, the rough mode decision, and templates (vvcb_cu_auto) the engine expands over the lists it has
// just produced -- every candidate the passes of EncCu::xCheckRDCostIntra (EL/EncCu.cpp:2453-2776) can reach: the final list x {DCT-II, transform skip},
// the final and the regular-only list x {DST-VII/DST-VII, DCT-II + LFNST 1, DCT-II + LFNST 2}.
void fetchFirstPass(CuCache& cu, const Snapshot& s, bool tsAllowed, bool mtsAllowed, bool mtsPass, bool lfnstWithMip)
{
  struct Tm { int lfnst, mts; uint8_t modes, skipMip; };
  std::vector<Tm> tm;
  tm.push_back(Tm{ 0, MTS_DCT2_DCT2, VVCB_AUTO_FINAL, 0 });
  if (tsAllowed) tm.push_back(Tm{ 0, MTS_SKIP, VVCB_AUTO_FINAL, 0 });
  if (mtsPass) tm.push_back(Tm{ 0, MTS_DST7_DST7, VVCB_AUTO_FINAL | VVCB_AUTO_REGULAR, 0 });
  tm.push_back(Tm{ 1, MTS_DCT2_DCT2, VVCB_AUTO_FINAL | VVCB_AUTO_REGULAR, (uint8_t)!lfnstWithMip });
  tm.push_back(Tm{ 2, MTS_DCT2_DCT2, VVCB_AUTO_FINAL | VVCB_AUTO_REGULAR, (uint8_t)!lfnstWithMip });
  std::vector<vvcb_cu_auto> autos(tm.size());
  for (size_t i = 0; i < tm.size(); i++) {
    memset(&autos[i], 0, sizeof(vvcb_cu_auto));
    makeJob(autos[i].job, cu, TuSpec{ 0, tm[i].lfnst, tm[i].mts, 0 }, s, tsAllowed, mtsAllowed, 0);
    autos[i].modes = tm[i].modes; autos[i].skip_mip = tm[i].skipMip;
  }
  const int maxAuto = 96, bs = cu.w * cu.h;
  // reused from CU to CU: fresh arrays of 96 blocks cost a walker ~0.5 GB of zero-filled, page-faulted memory per CTU
  static std::vector<int32_t> level;
  static std::vector<int16_t> reco, pred;
  if (level.size() < (size_t)maxAuto * bs) { level.resize((size_t)maxAuto * bs); reco.resize((size_t)maxAuto * bs); pred.resize((size_t)maxAuto * bs); }
  std::vector<vvcb_tu_result> res(maxAuto);
  std::vector<uint8_t> slots(maxAuto), tmpl(maxAuto);
  int nAuto = 0;
  vvcb_cu_request q;
  memset(&q, 0, sizeof(q));
  q.rects = cu.rects.data(); q.n_rects = (int)cu.rects.size(); q.rect_samples = cu.rectSamples.data(); q.n_rect_samples = cu.rectSamples.size();
  q.visit = &cu.visit; q.want_rmd = 1; q.result = &cu.res; q.detail = &cu.det;
  q.rates = &s.rates; q.states = &s.states;
  q.autos = autos.data(); q.n_autos = (int)autos.size(); q.max_auto = maxAuto; q.n_auto = &nAuto; q.auto_slot = slots.data(); q.auto_tmpl = tmpl.data();
  q.auto_level = level.data(); q.auto_reco = reco.data(); q.auto_pred = pred.data(); q.auto_results = res.data();
  const auto t0 = std::chrono::steady_clock::now();
  gpuCheck(vvcb_cu_eval(g_gpu, &q, 1), "vvcb_cu_eval:");
  g_st.engineNs += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  cu.pushed = true; cu.rmdValid = true;
  g_st.rmdRoundTrips++; g_st.visits++; g_st.jobsPrefetched += nAuto;
  for (int i = 0; i < nAuto; i++) {
    const Tm& t = tm[tmpl[i]];
    TuEntry& e = cu.tus[tuKey(slots[i], t.lfnst, t.mts)];
    e.job = autos[tmpl[i]].job;
    e.job.offset = (uint32_t)i * bs;
    e.job.intra_mode = t.lfnst ? (slots[i] >= VVCB_SLOT_MIP ? (uint8_t)PLANAR_IDX : (slots[i] < VVCB_SLOT_MRL1 ? slots[i] : cu.visit.mpm[1 + (slots[i] - VVCB_SLOT_MRL1) % 5])) : 0;
    e.res = res[i]; e.slot = slots[i];
    e.level.clear(); e.reco.clear();
    e.lv = level.data() + (size_t)i * bs; e.rc = reco.data() + (size_t)i * bs;    // valid until the next CU's prefetch, which starts a new cache
    e.rateHash = s.rateHash; e.stateHash = s.stateHash;
    std::vector<int16_t>& p = cu.pred[slots[i]];
    if (p.empty()) p.assign(pred.begin() + (size_t)i * bs, pred.begin() + (size_t)(i + 1) * bs);
  }
}

bool cuMatches(const CompArea& a) { return g_cu.valid && a.compID == COMPONENT_Y && a.x == g_cu.x && a.y == g_cu.y && (int)a.width == g_cu.w && (int)a.height == g_cu.h; }

bool tsAllowedFor(const CodingUnit& cu, int w, int h)
{
  const int maxSize = 1 << cu.cs->pps->getPpsRangeExtension().getLog2MaxTransformSkipBlockSize();    // TU::isTSAllowed, CL/UnitTools.cpp:4524
  return cu.cs->sps->getTransformSkipEnabledFlag() && !cu.transQuantBypass && !cu.ispMode && !cu.bdpcmMode && w <= maxSize && h <= maxSize && !cu.sbtInfo;
}
bool mtsAllowedFor(const CodingUnit& cu, int w, int h)
{
  return cu.cs->sps->getUseIntraMTS() && w <= MTS_INTRA_MAX_CU_SIZE && h <= MTS_INTRA_MAX_CU_SIZE && !cu.ispMode && !cu.sbtInfo && !cu.bdpcmMode;   // TU::isMTSAllowed, :4549
}

// the serve paths need the TU's candidate with the context the reference is in NOW; fetched on demand when the prefetch guessed otherwise
TuEntry& entryFor(const TransformUnit& tu, const Ctx& ctx, int mts, bool needStates)
{
  const CodingUnit& cu = *tu.cu;
  const PredictionUnit& pu = *tu.cs->getPU(tu.blocks[COMPONENT_Y].pos(), CHANNEL_TYPE_LUMA);
  const bool mip = PU::isMIP(pu, CHANNEL_TYPE_LUMA);
  const int slot = slotOf(pu, mip);
  if (!slotEvaluated(g_cu, slot)) die("full-RD candidate is not an evaluation slot of the visit");
  const bool ts = mts == MTS_SKIP;
  const int lfnst = ts ? 0 : cu.lfnstIdx;
  Snapshot fresh;
  const bool same = sameAsEntryCtx(ctx);
  if (!same) takeSnapshot(fresh, cu, ctx);
  const Snapshot& s = same ? g_entrySnap : fresh;
  TuSpec t = { slot, lfnst, mts, lfnst ? (mip ? (int)PLANAR_IDX : (int)PU::getFinalIntraMode(pu, CHANNEL_TYPE_LUMA)) : 0 };
  auto it = g_cu.tus.find(tuKey(slot, lfnst, mts));
  const bool hit = it != g_cu.tus.end() && it->second.rateHash == s.rateHash && (!needStates || it->second.stateHash == s.stateHash) && it->second.job.lambda == s.lambda &&
                   it->second.job.cbf_delta_bits == s.cbfDelta && it->second.job.qp_per == s.per[ts] && it->second.job.qp_rem == s.rem[ts] &&
                   (!lfnst || it->second.job.intra_mode == t.intraMode);
  if (!hit) {
    if (it != g_cu.tus.end()) { g_st.staleRate++; g_cu.tus.erase(it); }
    fetch(g_cu, false, std::vector<TuSpec>(1, t), s, tsAllowedFor(cu, g_cu.w, g_cu.h), mtsAllowedFor(cu, g_cu.w, g_cu.h), true);
    it = g_cu.tus.find(tuKey(slot, lfnst, mts));
  }
  return it->second;
}

Distortion shimSad(const DistParam& dp)
{
  if (!g_inRmd || g_curSlot < 0 || g_cu.det.sad[g_curSlot] == VVCB_SAT_NONE) die("SAD asked for a slot the visit did not evaluate");
  g_st.distServed++;
  return g_cu.det.sad[g_curSlot];
}
Distortion shimHad(const DistParam& dp)
{
  if (!g_inRmd || g_curSlot < 0 || g_cu.det.satd[g_curSlot] == VVCB_SAT_NONE) die("SATD asked for a slot the visit did not evaluate");
  g_st.distServed++;
  return g_cu.det.satd[g_curSlot];
}

bool tuServed(const TransformUnit& tu, ComponentID c)
{
  return g_inEst && c == COMPONENT_Y && !tu.noResidual && !tu.cu->ispMode && !tu.cu->bdpcmMode && CU::isIntra(*tu.cu) && cuMatches(tu.blocks[COMPONENT_Y]);
}

} // namespace

extern "C" {

bool __real__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib(IntraSearch*, CodingUnit&, Partitioner&, double, bool, int, int, bool);
void __real__ZN15IntraPrediction22initIntraPatternChTypeERK10CodingUnitRK8CompAreab(IntraPrediction*, const CodingUnit&, const CompArea&, bool);
void __real__ZN15IntraPrediction12initIntraMipERK14PredictionUnit(IntraPrediction*, const PredictionUnit&);
void __real__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit(IntraPrediction*, ComponentID, PelBuf&, const PredictionUnit&);
void __real__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit(IntraPrediction*, ComponentID, PelBuf&, const PredictionUnit&);
void __real__ZN6RdCost12setDistParamER9DistParamRK7AreaBufIKsES6_i11ComponentIDb(RdCost*, DistParam&, const CPelBuf&, const CPelBuf&, int, ComponentID, bool);
Distortion __real__ZN6RdCost11getDistPartERK7AreaBufIKsES4_i11ComponentID5DFuncPS3_(RdCost*, const CPelBuf&, const CPelBuf&, int, ComponentID, DFunc, const CPelBuf*);
void __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi(TrQuant*, TransformUnit&, const ComponentID&, const QpParam&, std::vector<TrMode>*, int);
void __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb(TrQuant*, TransformUnit&, const ComponentID&, const QpParam&, TCoeff&, const Ctx&, bool);
void __real__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam(TrQuant*, TransformUnit&, const ComponentID&, PelBuf&, const QpParam&);
void __real__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx(CABACWriter*, const TransformUnit&, ComponentID, CUCtx*);

// ---- IntraSearch::estIntraPredLumaQT, EL/IntraSearch.cpp:289 ------------------------------------------------------------------------
bool __wrap__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib(IntraSearch* is, CodingUnit& cu, Partitioner& pm, double best, bool mtsCheckRange, int mtsFirst, int mtsLast, bool moreProbFirst)
{
  const CodingStructure& cs = *cu.cs;
  const SPS& sps = *cs.sps;
  g_is = is;
  // configurations the engine does not cover run the reference's own code (none of them is in the shipped cfg)
  static const bool off = getenv("VVCB_SHIM_OFF") != nullptr;
  g_enabled = !off && cs.slice->getDepQuantEnabledFlag() && is->m_pcTrQuant->m_quant->m_useRDOQ && is->m_pcTrQuant->m_quant->m_useRDOQTS && !sps.getBDPCMEnabledFlag() && sps.getUseLFNST() &&
              !cs.pps->getPpsRangeExtension().getCrossComponentPredictionEnabledFlag() && !(cs.slice->getLmcsEnabledFlag() && is->m_pcReshape->getCTUFlag()) &&
              !is->m_pcEncCfg->getLumaLevelToDeltaQPMapping().isEnabled() && !cu.transQuantBypass && pm.chType == CHANNEL_TYPE_LUMA && is->m_pcEncCfg->getFastUDIUseMPMEnabled() &&
              is->m_pcEncCfg->getUseFastMIP() == true;
  if (!g_enabled) {
    static bool reg = false;
    if (!reg && !g_gpu) { reg = true; atexit(report); }
    Scope scEst(g_tEst);
    g_st.estCalls++;
    g_counting = true; g_countRmd = false;
    g_cntX = cu.firstPU->Y().x; g_cntY = cu.firstPU->Y().y; g_cntW = pm.currArea().lwidth(); g_cntH = pm.currArea().lheight();
    const bool r = __real__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib(is, cu, pm, best, mtsCheckRange, mtsFirst, mtsLast, moreProbFirst);
    g_counting = false;
    return r;
  }

  Scope scEst(g_tEst);
  ensureFrame(cs);
  g_st.estCalls++;
  std::unique_ptr<Scope> scPre(new Scope(g_tPre));
  const int w = pm.currArea().lwidth(), h = pm.currArea().lheight();
  PredictionUnit& pu = *cu.firstPU;
  const Position lt = pu.Y();

  // ---- the visit: geometry, availability (CL/IntraPrediction.cpp:1262-1267), MPMs, mode-bit prices, lambda ----
  vvcb_rmd_visit v;
  memset(&v, 0, sizeof(v));
  v.x = lt.x; v.y = lt.y; v.log2w = floorLog2(w); v.log2h = floorLog2(h);
  v.avail_al = unitAvail(cs, cu, lt.offset(-1, -1));
  int n;
  for (n = 0; n < w / 4 && unitAvail(cs, cu, lt.offset(4 * n, -1)); n++) {}
  v.n_above = n;
  for (n = 0; n < w / 4 && unitAvail(cs, cu, lt.offset(w + 4 * n, -1)); n++) {}
  v.n_above_right = n;
  for (n = 0; n < h / 4 && unitAvail(cs, cu, lt.offset(-1, 4 * n)); n++) {}
  v.n_left = n;
  for (n = 0; n < h / 4 && unitAvail(cs, cu, lt.offset(-1, h + 4 * n)); n++) {}
  v.n_below_left = n;
  unsigned mpm[NUM_MOST_PROBABLE_MODES];
  const int savedMrl = pu.multiRefIdx;
  pu.multiRefIdx = 0;
  v.num_mpm_cand = PU::getIntraMPMs(pu, mpm);
  pu.multiRefIdx = savedMrl;
  for (int i = 0; i < NUM_MOST_PROBABLE_MODES; i++) v.mpm[i] = mpm[i];
  if (!sps.getUseMIP()) v.flags |= VVCB_VISIT_NO_MIP;
  const auto& st = is->m_CABACEstimator->getCtx().m_CtxStore_Std;
  const unsigned mipCtx = DeriveCtx::CtxMipFlag(cu);
  v.rates.mip_flag[0] = st[Ctx::MipFlag(mipCtx)].estFracBits(0);      v.rates.mip_flag[1] = st[Ctx::MipFlag(mipCtx)].estFracBits(1);
  v.rates.mrl_bin0[0] = st[Ctx::MultiRefLineIdx(0)].estFracBits(0);   v.rates.mrl_bin0[1] = st[Ctx::MultiRefLineIdx(0)].estFracBits(1);
  v.rates.mrl_bin1[0] = st[Ctx::MultiRefLineIdx(1)].estFracBits(0);   v.rates.mrl_bin1[1] = st[Ctx::MultiRefLineIdx(1)].estFracBits(1);
  v.rates.isp_bin0_0  = st[Ctx::ISPMode(0)].estFracBits(0);
  v.rates.mpm_flag[0] = st[Ctx::IntraLumaMpmFlag()].estFracBits(0);   v.rates.mpm_flag[1] = st[Ctx::IntraLumaMpmFlag()].estFracBits(1);
  v.rates.planar_flag[0] = st[Ctx::IntraLumaPlanarFlag(1)].estFracBits(0); v.rates.planar_flag[1] = st[Ctx::IntraLumaPlanarFlag(1)].estFracBits(1);
  v.sqrt_lambda = is->m_pcRdCost->getMotionLambda(cu.transQuantBypass) * FRAC_BITS_SCALE;

  // ---- the reconstructed neighbourhood the reference lines come from: 4 rows above over 2w + 8 columns, 4 columns left over 2h + 4 rows ----
  std::vector<vvcb_rect> rects;
  std::vector<int16_t> samples;
  {
    const CPelBuf reco = cs.picture->getRecoBuf(COMPONENT_Y);
    const int pw = cs.picture->lwidth(), ph = cs.picture->lheight();
    auto add = [&](int x0, int y0, int rw, int rh) {
      vvcb_rect r = { (int16_t)x0, (int16_t)y0, (int16_t)rw, (int16_t)rh, (uint32_t)samples.size() };
      for (int yy = 0; yy < rh; yy++) { const Pel* src = reco.bufAt(x0, y0 + yy); samples.insert(samples.end(), src, src + rw); }
      rects.push_back(r);
    };
    if (lt.y >= 4) { const int x0 = std::max(0, lt.x - 4), x1 = std::min(pw, lt.x + 2 * w + 4); add(x0, lt.y - 4, x1 - x0, 4); }
    if (lt.x >= 4) { const int y1 = std::min(ph, lt.y + 2 * h + 4); add(lt.x - 4, lt.y, 4, y1 - lt.y); }
  }
  uint64_t nbh = mix(samples.data(), samples.size() * sizeof(int16_t));
  nbh = mix(&v, sizeof(v), nbh);                                       // availability, MPMs, mode-bit prices and lambda are part of the visit's identity

  if (g_cu.valid && g_cu.x == lt.x && g_cu.y == lt.y && g_cu.w == w && g_cu.h == h && g_cu.nbhHash == nbh) g_st.cuReuse++;
  else {
    g_cu = CuCache();
    g_cu.valid = true; g_cu.x = lt.x; g_cu.y = lt.y; g_cu.w = w; g_cu.h = h; g_cu.nbhHash = nbh;
    g_cu.visit = v; g_cu.rects.swap(rects); g_cu.rectSamples.swap(samples);
  }

  // ---- which part of the function runs (EL/IntraSearch.cpp:312-345, 430) ----
  const bool lfnstLoad = sps.getUseLFNST() && cu.lfnstIdx != 0;
  int mtsUsage = 0;
  if (w <= MTS_INTRA_MAX_CU_SIZE && h <= MTS_INTRA_MAX_CU_SIZE && sps.getUseIntraMTS()) mtsUsage = (sps.getUseLFNST() && cu.mtsFlag == 1) ? 2 : 1;
  if (w * h < 64 && !is->m_pcEncCfg->getUseFastLFNST()) mtsUsage = 0;
  const bool rmdRuns = mtsUsage != 2 && !lfnstLoad;

  rememberEntryCtx(cu, is->m_CABACEstimator->getCtx());
  const Snapshot& snap = g_entrySnap;
  const bool tsAllowed = tsAllowedFor(cu, w, h), mtsAllowed = mtsAllowedFor(cu, w, h);
  static const bool noPrefetch = getenv("VVCB_SHIM_NO_PREFETCH") != nullptr;
  static const bool noMega = getenv("VVCB_SHIM_NO_MEGA") != nullptr;           // candidates fetched per estIntraPredLumaQT call instead of per CU
  static const bool twoTrips = getenv("VVCB_SHIM_TWO_TRIPS") != nullptr;       // lists first, candidates in a second round trip (the shim names them)
  if (rmdRuns && !g_cu.rmdValid) {
    if (!noPrefetch && !noMega && !twoTrips && cu.lfnstIdx == 0 && cu.mtsFlag == 0) fetchFirstPass(g_cu, snap, tsAllowed, mtsAllowed, mtsUsage == 1, allowLfnstWithMip(Size(w, h)));
    else fetch(g_cu, true, std::vector<TuSpec>(), snap, tsAllowed, mtsAllowed, false);
  }

  // ---- the candidates the full-RD loop of this pass can reach (:1158; transforms per xRecurIntraCodingLumaQT :3340-3501) ----
  if (!noPrefetch) {
    struct M { bool mip; int mrl; int mode; };
    std::vector<M> modes;
    const bool mipWithLfnst = cu.lfnstIdx == 0 || allowLfnstWithMip(Size(w, h));
    auto addMode = [&](bool mip, int mrl, int mode) {
      if (mip && !mipWithLfnst) return;
      for (const M& m : modes) if (m.mip == mip && m.mrl == mrl && m.mode == mode) return;
      modes.push_back(M{ mip, mrl, mode });
    };
    if (rmdRuns) for (int i = 0; i < g_cu.res.n_final; i++) addMode(g_cu.res.final_mode[i].mip != 0, g_cu.res.final_mode[i].mrl, g_cu.res.final_mode[i].mode);
    else if (mtsUsage == 2) { for (int i = 0; i < is->m_savedNumRdModes[cu.lfnstIdx]; i++) { const auto& m = is->m_savedRdModeList[cu.lfnstIdx][i]; if (m.ispMod == NOT_INTRA_SUBPARTITIONS) addMode(m.mipFlg, m.mRefId, m.modeId); } }
    else {                                                             // LFNST pass: the saved lists + the MPMs appended again (:777-802)
      for (const auto& m : is->m_uiSavedRdModeListLFNST) if (m.ispMod == NOT_INTRA_SUBPARTITIONS) addMode(m.mipFlg, m.mRefId, m.modeId);
      for (int j = 0; j < v.num_mpm_cand; j++) addMode(false, 0, v.mpm[j]);
    }
    std::vector<TuSpec> specs;
    for (const M& m : modes) {
      const int slot = slotOfMode(g_cu, m.mip, m.mrl, m.mode);
      if (!slotEvaluated(g_cu, slot)) continue;
      const int intraMode = m.mip ? (int)PLANAR_IDX : m.mode;         // PU::getFinalIntraMode of a luma block is its own mode
      if (cu.mtsFlag) {
        const int first = mtsCheckRange ? mtsFirst : 0, last = mtsCheckRange ? mtsLast : 3;
        for (int ti = first; ti <= last; ti++) {
          int mts = MTS_DST7_DST7 + ti;
          if (moreProbFirst && ti == 1) mts = m.mode < 34 ? MTS_DST7_DCT8 : MTS_DCT8_DST7;
          if (moreProbFirst && ti == 2) mts = m.mode < 34 ? MTS_DCT8_DST7 : MTS_DST7_DCT8;
          specs.push_back(TuSpec{ slot, 0, mts, 0 });
        }
      } else {
        specs.push_back(TuSpec{ slot, cu.lfnstIdx, MTS_DCT2_DCT2, cu.lfnstIdx ? intraMode : 0 });
        if (tsAllowed && !cu.lfnstIdx) specs.push_back(TuSpec{ slot, 0, MTS_SKIP, 0 });
      }
    }
    // First pass of a CU: EncCu::xCheckRDCostIntra (EL/EncCu.cpp:2453-2776) goes on to call this function for (lfnst 0, MTS index 0),
    // (lfnst 1) and (lfnst 2) with lists that derive from this pass's (saved lists + MPMs); their candidates ride along so that those
    // calls find everything cached.  VVCB_SHIM_NO_MEGA=1 fetches per call instead.
    if (rmdRuns && !noMega && cu.lfnstIdx == 0 && cu.mtsFlag == 0) {
      std::vector<M> later = modes;
      auto addLater = [&](bool mip, int mrl, int mode) {
        for (const M& m : later) if (m.mip == mip && m.mrl == mrl && m.mode == mode) return;
        later.push_back(M{ mip, mrl, mode });
      };
      for (int i = 0; i < g_cu.det.n_reg; i++) addLater(g_cu.det.reg_mode[i].mip != 0, g_cu.det.reg_mode[i].mrl, g_cu.det.reg_mode[i].mode);
      const bool lfnstMip = allowLfnstWithMip(Size(w, h));
      for (const M& m : later) {
        const int slot = slotOfMode(g_cu, m.mip, m.mrl, m.mode);
        if (!slotEvaluated(g_cu, slot)) continue;
        const int intraMode = m.mip ? (int)PLANAR_IDX : m.mode;
        if (mtsUsage == 1) specs.push_back(TuSpec{ slot, 0, MTS_DST7_DST7, 0 });
        if (!m.mip || lfnstMip) { specs.push_back(TuSpec{ slot, 1, MTS_DCT2_DCT2, intraMode }); specs.push_back(TuSpec{ slot, 2, MTS_DCT2_DCT2, intraMode }); }
      }
    }
    fetch(g_cu, false, specs, snap, tsAllowed, mtsAllowed, false);
  }

  scPre.reset();
  g_inEst = true; g_inRmd = false; g_curSlot = -1; g_pend = PendingTu();
  const bool ret = __real__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib(is, cu, pm, best, mtsCheckRange, mtsFirst, mtsLast, moreProbFirst);
  g_inEst = false; g_inRmd = false; g_pend = PendingTu();
  return ret;
}

// ---- reference-sample fetch: the engine builds the lines from the reconstruction rectangles ------------------------------------------
void __wrap__ZN15IntraPrediction22initIntraPatternChTypeERK10CodingUnitRK8CompAreab(IntraPrediction* ip, const CodingUnit& cu, const CompArea& area, bool forceRefFilterFlag)
{
  if (g_inEst && !cu.ispMode && !cu.bdpcmMode && cuMatches(area)) {
    g_inRmd = forceRefFilterFlag;                                      // only the RMD block passes true (EL/IntraSearch.cpp:483, :645)
    g_st.refFetchSkipped++;
    return;
  }
  if (countArea(area) && !cu.ispMode) g_countRmd = forceRefFilterFlag;
  if (g_profile) { Scope sc(g_prof[0][catOf(area.compID, cu)]); __real__ZN15IntraPrediction22initIntraPatternChTypeERK10CodingUnitRK8CompAreab(ip, cu, area, forceRefFilterFlag); return; }
  __real__ZN15IntraPrediction22initIntraPatternChTypeERK10CodingUnitRK8CompAreab(ip, cu, area, forceRefFilterFlag);
}

void __wrap__ZN15IntraPrediction12initIntraMipERK14PredictionUnit(IntraPrediction* ip, const PredictionUnit& pu)
{
  if (g_inEst && cuMatches(pu.Y()) && !pu.cu->ispMode) { g_inRmd = true; return; }    // :712-714: the MIP pass of the RMD block
  if (countArea(pu.Y())) g_countRmd = true;
  __real__ZN15IntraPrediction12initIntraMipERK14PredictionUnit(ip, pu);
}

static void servePrediction(PelBuf& pred, const PredictionUnit& pu, bool mip)
{
  WrapScope sc;
  const int slot = slotOf(pu, mip);
  if (!slotEvaluated(g_cu, slot)) die("prediction asked for a mode that is not an evaluation slot of the visit");
  if (g_inRmd) { g_curSlot = slot; g_st.predSkipped++; return; }        // only its SAD / SATD are read (setDistParam wrapper)
  auto it = g_cu.pred.find(slot);
  if (it == g_cu.pred.end()) {                                          // not prefetched: fetch the candidate the loop is about to code
    Snapshot fresh;
    const bool same = sameAsEntryCtx(g_is->m_CABACEstimator->getCtx());
    if (!same) takeSnapshot(fresh, *pu.cu, g_is->m_CABACEstimator->getCtx());
    const Snapshot& s = same ? g_entrySnap : fresh;
    const int lfnst = pu.cu->lfnstIdx;
    int mts = MTS_DCT2_DCT2;
    TuSpec t = { slot, lfnst, mts, lfnst ? (mip ? (int)PLANAR_IDX : (int)pu.intraDir[0]) : 0 };
    fetch(g_cu, false, std::vector<TuSpec>(1, t), s, tsAllowedFor(*pu.cu, g_cu.w, g_cu.h), mtsAllowedFor(*pu.cu, g_cu.w, g_cu.h), true);
    it = g_cu.pred.find(slot);
  }
  if ((int)pred.width != g_cu.w || (int)pred.height != g_cu.h) die("prediction block size differs from the visit");
  const int16_t* src = it->second.data();
  for (int y = 0; y < g_cu.h; y++) memcpy(pred.bufAt(0, y), src + (size_t)y * g_cu.w, g_cu.w * sizeof(int16_t));
  g_st.predServed++;
}

void __wrap__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit(IntraPrediction* ip, ComponentID c, PelBuf& pred, const PredictionUnit& pu)
{
  if (g_inEst && c == COMPONENT_Y && !pu.cu->ispMode && !pu.cu->bdpcmMode && cuMatches(pu.Y())) { servePrediction(pred, pu, false); return; }
  if (c == COMPONENT_Y && g_countRmd && countArea(pu.Y()) && !pu.cu->ispMode) g_satdEvals++;
  if (g_profile) { Scope sc(g_prof[1][catOf(c, *pu.cu)]); __real__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit(ip, c, pred, pu); return; }
  __real__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit(ip, c, pred, pu);
}

void __wrap__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit(IntraPrediction* ip, ComponentID c, PelBuf& pred, const PredictionUnit& pu)
{
  if (g_inEst && c == COMPONENT_Y && !pu.cu->ispMode && cuMatches(pu.Y())) { servePrediction(pred, pu, true); return; }
  if (c == COMPONENT_Y && g_countRmd && countArea(pu.Y()) && !pu.cu->ispMode) g_satdEvals++;
  if (g_profile) { Scope sc(g_prof[1][catOf(c, *pu.cu)]); __real__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit(ip, c, pred, pu); return; }
  __real__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit(ip, c, pred, pu);
}

// ---- SAD / SATD of the RMD block (seam S1): distParam.distFunc( distParam ), EL/IntraSearch.cpp:515 -------------------------------------
void __wrap__ZN6RdCost12setDistParamER9DistParamRK7AreaBufIKsES6_i11ComponentIDb(RdCost* rc, DistParam& dp, const CPelBuf& org, const CPelBuf& cur, int bitDepth, ComponentID c, bool useHadamard)
{
  __real__ZN6RdCost12setDistParamER9DistParamRK7AreaBufIKsES6_i11ComponentIDb(rc, dp, org, cur, bitDepth, c, useHadamard);
  if (g_inEst && c == COMPONENT_Y && g_cu.rmdValid && (int)org.width == g_cu.w && (int)org.height == g_cu.h) dp.distFunc = useHadamard ? shimHad : shimSad;
}

// ---- TrQuant::transformNxN( trModes ), CL/TrQuant.cpp:1049: pre-selection among DCT-II and transform skip ------------------------------------
void __wrap__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi(TrQuant* tq, TransformUnit& tu, const ComponentID& c, const QpParam& qp, std::vector<TrMode>* modes, int maxCand)
{
  g_pend = PendingTu();
  if (!tuServed(tu, c) || tu.cu->lfnstIdx != 0 || modes->empty()) {
    std::unique_ptr<Scope> sc(g_profile ? new Scope(g_prof[3][catOf(c, *tu.cu)]) : nullptr);
    __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi(tq, tu, c, qp, modes, maxCand);
    return;
  }
  WrapScope sc;
  const int k = (int)modes->size();
  std::vector<int32_t> sums(k);
  std::vector<uint8_t> sel(k);
  for (int i = 0; i < k; i++) sums[i] = entryFor(tu, g_is->m_CABACEstimator->getCtx(), (*modes)[i].first, false).res.abs_sum_coeff;
  vvcb_mts_preselect(sums.data(), k, g_cu.w, g_cu.h, maxCand, sel.data());
  for (int i = 0; i < k; i++) (*modes)[i].second = sel[i] != 0;
  tu.mtsIdx = (*modes)[k - 1].first;                                   // what the reference's loop leaves behind (the caller sets it again, EL/IntraSearch.cpp:2966)
  g_st.preselServed++;
}

// ---- TrQuant::transformNxN( quant ), CL/TrQuant.cpp:1127: levels and absSum ---------------------------------------------------------------
void __wrap__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb(TrQuant* tq, TransformUnit& tu, const ComponentID& c, const QpParam& qp, TCoeff& absSum, const Ctx& ctx, bool loadTr)
{
  g_pend = PendingTu();
  if (!tuServed(tu, c)) {
    g_countRmd = false;                                  // the full-RD loop has begun
    g_tuQuantReal++;
    std::unique_ptr<Scope> sc(g_profile ? new Scope(g_prof[4][catOf(c, *tu.cu)]) : nullptr);
    __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb(tq, tu, c, qp, absSum, ctx, loadTr);
    return;
  }
  WrapScope sc;
  const bool ts = tu.mtsIdx == MTS_SKIP;
  TuEntry& e = entryFor(tu, ctx, tu.mtsIdx, false);
  if (e.job.qp_per != qp.per(ts) || e.job.qp_rem != qp.rem(ts) || e.job.lambda != tq->m_quant->getLambda()) die("QP / lambda of the TU differ from the ones its candidate was computed with");
  CoeffBuf lv = tu.getCoeffs(c);
  for (int y = 0; y < g_cu.h; y++) memcpy(&lv.at(0, y), &e.lv[(size_t)y * g_cu.w], g_cu.w * sizeof(TCoeff));
  absSum = e.res.abs_sum_level;
  TU::setCbfAtDepth(tu, c, tu.depth, absSum > 0);
  g_pend.tu = &tu; g_pend.e = &e; g_pend.recoBuf = tu.cs->getRecoBuf(tu.blocks[c]).buf;
  g_st.quantServed++; g_st.quantDq += !ts; g_st.quantTs += ts; g_st.quantLfnst += e.job.lfnst_idx != 0;
}

// ---- TrQuant::invTransformNxN, CL/TrQuant.cpp:561: the residual that reconstructs to the engine's block --------------------------------------
void __wrap__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam(TrQuant* tq, TransformUnit& tu, const ComponentID& c, PelBuf& resi, const QpParam& qp)
{
  if (c != COMPONENT_Y || g_pend.tu != &tu || !g_pend.e) {
    std::unique_ptr<Scope> sc(g_profile ? new Scope(g_prof[5][catOf(c, *tu.cu)]) : nullptr);
    __real__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam(tq, tu, c, resi, qp);
    return;
  }
  WrapScope sc;
  const std::vector<int16_t>& pred = g_cu.pred[g_pend.e->slot];
  for (int y = 0; y < g_cu.h; y++)
    for (int x = 0; x < g_cu.w; x++) resi.at(x, y) = g_pend.e->rc[(size_t)y * g_cu.w + x] - pred[(size_t)y * g_cu.w + x];   // PelBuf::reconstruct clips pred + resi back to reco
  g_st.invServed++;
}

// ---- RdCost::getDistPart( org, reco, DF_SSE ), EL/IntraSearch.cpp:3160 ----------------------------------------------------------------------
Distortion __wrap__ZN6RdCost11getDistPartERK7AreaBufIKsES4_i11ComponentID5DFuncPS3_(RdCost* rc, const CPelBuf& org, const CPelBuf& cur, int bitDepth, ComponentID c, DFunc f, const CPelBuf* orgLuma)
{
  if (g_inEst && g_pend.e && c == COMPONENT_Y && f == DF_SSE && cur.buf == g_pend.recoBuf && (int)cur.width == g_cu.w && (int)cur.height == g_cu.h) {
    g_st.sseServed++;
    return g_pend.e->res.sse;
  }
  std::unique_ptr<Scope> sc(g_profile ? new Scope(g_prof[6][c != COMPONENT_Y ? 2 : 0]) : nullptr);
  return __real__ZN6RdCost11getDistPartERK7AreaBufIKsES4_i11ComponentID5DFuncPS3_(rc, org, cur, bitDepth, c, f, orgLuma);
}

// ---- CABACWriter::residual_coding on the bit estimator (EL/CABACWriter.cpp:3773, from IntraSearch::xEncCoeffQT) ----------------------------------
// The estimator state after the call is dropped by every caller inside estIntraPredLumaQT (restored to ctxStart / ctxBest copies that are
// themselves restored, EL/IntraSearch.cpp:1222, :1377, :3448), so only the bit count matters here; the final cu_residual of
// EncCu::xCheckRDCostIntra calls residual_coding inside CABACWriter.cpp and is not intercepted.
void __wrap__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx(CABACWriter* cw, const TransformUnit& tu, ComponentID c, CUCtx* cuCtx)
{
  if (g_inEst && c == COMPONENT_Y && g_pend.tu == &tu && g_pend.e && !cw->m_BinEncoder.isEncoding() && cw == g_is->m_CABACEstimator) {
    WrapScope sc;
    uint64_t stateHash = g_entrySnap.stateHash;
    if (!sameAsEntryCtx(cw->getCtx())) { vvcb_ctx_states st; fillStates(st, cw->getCtx()); stateHash = mix(&st, sizeof(st)); }
    if (stateHash == g_pend.e->stateHash) {
      BitEstimatorBase* be = dynamic_cast<BitEstimatorBase*>(&cw->m_BinEncoder);
      if (!be) die("the estimator is not a BitEstimator");
      be->m_EstFracBits += g_pend.e->res.frac_bits;
      g_st.bitsServed++;
      return;
    }
  }
  if (g_inEst && c == COMPONENT_Y && !tu.cu->ispMode) g_st.bitsReal++;
  std::unique_ptr<Scope> sc(g_profile ? new Scope(g_prof[7][catOf(c, *tu.cu)]) : nullptr);
  __real__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx(cw, tu, c, cuCtx);
}

} // extern "C"
