// TEST INFRASTRUCTURE -- executes the real RMD kernels (vvc_intra_b200/csrc/vvcb_rmd.cuh) on host threads
// through cuda_emul.h and exposes one C entry point for tests/test_kernel_emulation.py.
#include "cuda_emul.h"
#include <string.h>
#include <stdlib.h>
#include "../../vvc_intra_b200/csrc/vvcb_rmd.cuh"
#include "../../vvc_intra_b200/csrc/vvcb_tu.cuh"
#include "../../vvc_intra_b200/csrc/vvcb_feat.cuh"
#include "../../vvc_intra_b200/csrc/vvcb_dq.cuh"
#include "../../vvc_intra_b200/csrc/vvcb_rate.cuh"
#include <vector>
#include "../../vvc_intra_b200/csrc/vvcb_romfill.h"

thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local EmuBlock* emuBlock;

struct ThreadArg { unsigned tid, bid, block, grid; EmuBlock* blk; const std::function<void()>* fn; };

static void* thread_main(void* p)
{
  ThreadArg* a = (ThreadArg*)p;
  threadIdx = { a->tid, 0, 0 }; blockIdx = { a->bid, 0, 0 }; blockDim = { a->block, 1, 1 }; gridDim = { a->grid, 1, 1 };
  emuBlock = a->blk;
  (*a->fn)();
  return nullptr;
}

void emu_launch(unsigned grid, unsigned block, const std::function<void()>& fn)
{
  for (unsigned b = 0; b < grid; b++) {
    EmuBlock blk;
    pthread_barrier_init(&blk.bar, nullptr, block);
    blk.warps.resize((block + 31) / 32);
    for (unsigned w = 0; w < blk.warps.size(); w++) {
      const unsigned cnt = (w + 1) * 32 <= block ? 32 : block - w * 32;
      pthread_barrier_init(&blk.warps[w].bar, nullptr, cnt);
    }
    std::vector<pthread_t> th(block);
    std::vector<ThreadArg> args(block);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 1 << 20);
    for (unsigned t = 0; t < block; t++) {
      args[t] = { t, b, block, grid, &blk, &fn };
      pthread_create(&th[t], &attr, thread_main, &args[t]);
    }
    for (unsigned t = 0; t < block; t++) pthread_join(th[t], nullptr);
    pthread_attr_destroy(&attr);
  }
}

template <int TILE, int KIND, int MODE> static void emul_eval_bucket(const EvalParams& P)
{
  emu_launch(1, EvalCfg<MODE == 1>::kThreads, [&] { rmd_eval_kernel<TILE, KIND, MODE>(P); });
}

extern "C" int emul_rmd_eval(const int16_t* orig, const int16_t* reco, int stride, int bd, int ctu,
                             const vvcb_rmd_visit* visits, int n, vvcb_rmd_result* results, vvcb_rmd_detail* details,
                             int16_t* predOut)
{
  static Rom rom;
  fill_rom(rom);
  std::vector<WorkItem> items((size_t)n * 60 + 8);
  PlanState plan;
  memset(&plan, 0, sizeof(plan));
  const int pack = predOut ? 0 : 1;
  if (!predOut && n <= 96) {                  // walk-sized batch, as launch_rmd runs it (the plan state is written whole: start from garbage)
    memset(&plan, 0xA5, sizeof(plan));
    emu_launch(1, 256, [&] { rmd_plan_small(visits, n, ctu, &plan, items.data(), pack); });
  } else {
    emu_launch((n + 255) / 256, 256, [&] { rmd_plan_count(visits, n, ctu, &plan, pack); });
    emu_launch(1, 32, [&] { rmd_plan_scan(&plan); });
    emu_launch((n + 255) / 256, 256, [&] { rmd_plan_fill(visits, n, ctu, &plan, items.data(), pack); });
  }
  EvalParams P;
  P.visits = visits; P.items = items.data(); P.plan = &plan;
  std::vector<uint32_t> sm((size_t)2 * VVCB_NUM_SLOTS * n);
  P.sadSM = sm.data(); P.satdSM = (details || predOut) ? sm.data() + (size_t)VVCB_NUM_SLOTS * n : nullptr; P.nVisits = n;   // as launch_rmd (vvcb_api.cu)
  P.orig = orig; P.reco = reco; P.stride = stride; P.bd = bd; P.ctu = ctu; P.rom = &rom; P.predOut = predOut;
  if (!predOut && n <= 96) {
    // walk-sized batch, as launch_rmd (vvcb_api.cu) runs it: one any-bucket launch for the packed items, one for the plain ones
    for (int mode = 1; mode >= 0; mode--) {
      EvalAny A;
      A.n = 0;
      int cta = 0;
      for (int b = 0; b < kNumBuckets; b++) {
        bool occupied = mode == 0 ? plan.count[b] != 0 : false;
        if (mode == 1) for (int i = 0; i < pack_shape_count(b / kNumKinds); i++) occupied = occupied || plan.count[kNumBuckets + pack_shape(b / kNumKinds, i) * kNumKinds + b % kNumKinds];
        if (!occupied) continue;
        A.bucket[A.n] = (unsigned char)b; A.firstCta[A.n] = cta; cta += 1 + (b & 1); A.n++;          // one or two CTAs per bucket: both walk the bucket's cursor
      }
      A.firstCta[A.n] = cta;
      if (!A.n) continue;
      if (mode) emu_launch(cta, EvalCfg<true>::kThreads, [&] { rmd_eval_any_kernel<1>(P, A); });
      else      emu_launch(cta, EvalCfg<false>::kThreads, [&] { rmd_eval_any_kernel<0>(P, A); });
    }
  } else
  for (int b = 0; b < kNumBuckets; b++) {
    bool packed = false;
    for (int i = 0; i < pack_shape_count(b / kNumKinds); i++) packed = packed || plan.count[kNumBuckets + pack_shape(b / kNumKinds, i) * kNumKinds + b % kNumKinds];
    if (packed) VVCB_FOR_BUCKET(b, 1, emul_eval_bucket, P);
    if (plan.count[b] && predOut) VVCB_FOR_BUCKET(b, 2, emul_eval_bucket, P);
    if (plan.count[b] && !predOut) VVCB_FOR_BUCKET(b, 0, emul_eval_bucket, P);
  }
  if (details) emu_launch((n + 31) / 32, 256, [&] { rmd_detail_kernel(visits, n, ctu, details, P.sadSM, P.satdSM); });
  emu_launch((n + kListThreads - 1) / kListThreads, kListThreads, [&] { rmd_lists_kernel(visits, n, ctu, results, details, P.sadSM, P.satdSM, nullptr); });
  return 0;
}

// the list kernel alone, writing brief records (vvcb_rmd_eval_brief), from given SAD / SATD tables
extern "C" int emul_rmd_brief(const vvcb_rmd_visit* visits, int n, int ctu, const vvcb_rmd_detail* details, vvcb_rmd_brief* brief)
{
  std::vector<uint32_t> sm((size_t)2 * VVCB_NUM_SLOTS * n);
  for (int i = 0; i < n; i++)
    for (int s = 0; s < VVCB_NUM_SLOTS; s++) { sm[scratch_at(s, (unsigned)i, n)] = details[i].sad[s]; sm[(size_t)VVCB_NUM_SLOTS * n + scratch_at(s, (unsigned)i, n)] = details[i].satd[s]; }
  emu_launch((n + kListThreads - 1) / kListThreads, kListThreads, [&] { rmd_lists_kernel(visits, n, ctu, nullptr, nullptr, sm.data(), sm.data() + (size_t)VVCB_NUM_SLOTS * n, brief); });
  return 0;
}

// the three launches of vvcb_tu_eval (vvcb_api.cu): transform pass, dependent quantisation, reconstruction pass
extern "C" int emul_tu_eval(const int16_t* orig, int stride, int bd, const vvcb_tu_job* jobs, int n, const int16_t* resi, const int16_t* pred,
                            size_t nSamples, const vvcb_dq_rates* rates, const vvcb_ctx_states* states, int nRates,
                            int32_t* coeff, int32_t* level, int16_t* reco, vvcb_tu_result* results)
{
  static TrRom rom;
  static DqRom dqRom;
  fill_tr_rom(rom);
  fill_dq_rom(dqRom);
  std::vector<int> order, tsOrder;
  for (int i = 0; i < n; i++) {
    if ((jobs[i].flags & (VVCB_TU_QUANT | VVCB_TU_DEPQUANT)) == (VVCB_TU_QUANT | VVCB_TU_DEPQUANT)) order.push_back(i);
    else if ((jobs[i].flags & (VVCB_TU_QUANT | VVCB_TU_RDOQ_TS)) == (VVCB_TU_QUANT | VVCB_TU_RDOQ_TS)) tsOrder.push_back(i);
  }
  const int nDq = (int)order.size();
  std::vector<int32_t> dqCoeff(nSamples), dqDeq(nSamples, 0);
  memset(level, 0, nSamples * sizeof(int32_t));
  TuParams P;
  std::vector<int> smallList, largeList;
  for (int i = 0; i < n; i++) (jobs[i].log2w + jobs[i].log2h <= 8 ? smallList : largeList).push_back(i);
  auto launch_tu = [&](TuParams Q) {
    if (!smallList.empty()) { Q.list = smallList.data(); Q.n = (int)smallList.size(); emu_launch(2, kTuThreads, [&] { tu_eval_kernel<32>(Q); }); }
    if (!largeList.empty()) { Q.list = largeList.data(); Q.n = (int)largeList.size(); emu_launch(2, kTuThreads, [&] { tu_eval_kernel<128>(Q); }); }
  };
  P.jobs = jobs; P.list = nullptr; P.n = 0; P.resi = resi; P.pred = pred; P.coeff = coeff; P.level = level; P.reco = reco; P.results = results;
  P.orig = orig; P.stride = stride; P.bd = bd; P.rom = &rom;
  P.dqCoeff = dqCoeff.data(); P.dqDeq = dqDeq.data(); P.phase = 0;
  launch_tu(P);
  if (nDq) {
    std::vector<DqRateTab> tabs(nRates);
    emu_launch(nRates, 32, [&] { dq_rate_kernel(rates, nRates, tabs.data()); });
    const int grid = 2;
    std::vector<uint8_t> scratch((size_t)grid * kDqGroups * kDqSlotBytes);
    std::vector<int> firstRaw(nDq), orderSorted(nDq), firstSorted(nDq);
    emu_launch(2, kDqThreads, [&] { dq_first_kernel(jobs, order.data(), nDq, dqCoeff.data(), &dqRom, bd, firstRaw.data()); });
    std::vector<int> binCount(kDqBins, 0);
    const int sortGrid = (nDq + kDqSortPerBlock - 1) / kDqSortPerBlock;
    const bool sortJobs = nDq > 48;            // the product's rule (vvcb_api.cu: more than 2048 jobs) at the scale of the test sets: both paths run
    if (sortJobs) {
      emu_launch(sortGrid, kDqSortThreads, [&] { dq_hist_kernel(firstRaw.data(), nDq, binCount.data()); });
      emu_launch(1, 32, [&] { dq_scan_kernel(binCount.data()); });
      emu_launch(sortGrid, kDqSortThreads, [&] { dq_scatter_kernel(order.data(), firstRaw.data(), nDq, binCount.data(), orderSorted.data(), firstSorted.data()); });
    }
    DqParams D;
    D.jobs = jobs; D.order = sortJobs ? orderSorted.data() : order.data(); D.firstPos = sortJobs ? firstSorted.data() : firstRaw.data(); D.n = nDq;
    D.coeff = dqCoeff.data(); D.level = level; D.deq = dqDeq.data(); D.results = results;
    D.rates = rates; D.tabs = tabs.data(); D.rom = &dqRom; D.scratch = scratch.data(); D.bd = bd; D.sparse = !sortJobs;
    emu_launch(grid, kDqThreads, [&] { dq_kernel(D); });
  }
  if (!tsOrder.empty()) {
    RdoqParams R;
    R.jobs = jobs; R.order = tsOrder.data(); R.n = (int)tsOrder.size(); R.coeff = dqCoeff.data(); R.level = level; R.deq = dqDeq.data();
    R.results = results; R.rates = rates; R.rom = &dqRom; R.bd = bd;
    emu_launch(2, kTsThreads, [&] { rdoq_ts_kernel(R); });
  }
  if (nDq || !tsOrder.empty()) {
    P.phase = 1;
    launch_tu(P);
  }
  std::vector<int> rateOrder;
  for (int i = 0; i < n; i++) if ((jobs[i].flags & (VVCB_TU_QUANT | VVCB_TU_RATE)) == (VVCB_TU_QUANT | VVCB_TU_RATE)) rateOrder.push_back(i);
  if (!rateOrder.empty()) {
    static RateRom rr;
    for (int i = 0; i < 512; i++) rr.binFracBits[i] = kBinFracBits[i];
    RateParams R;
    R.jobs = jobs; R.order = rateOrder.data(); R.n = (int)rateOrder.size(); R.level = level; R.results = results; R.states = states;
    R.rom = &dqRom; R.rate = &rr; R.depQuant = 1;
    emu_launch(2, kRateThreads, [&] { rate_kernel(R); });
  }
  return 0;
}

extern "C" int emul_features_eval(const int16_t* orig, int stride, const vvcb_feat_job* jobs, int n, vvcb_feat_result* results)
{
  FeatParams P;
  P.jobs = jobs; P.n = n; P.results = results; P.orig = orig; P.stride = stride;
  emu_launch(3, kFeatThreads, [&] { features_kernel(P); });
  return 0;
}

extern "C" int emul_ctu_hads(const int16_t* orig, int stride, int width, int height, int ctu, int32_t* out)
{
  const int perRow = (width + ctu - 1) / ctu, rows = (height + ctu - 1) / ctu;
  memset(out, 0, sizeof(int32_t) * perRow * rows);
  emu_launch(2, 256, [&] { ctu_hads_kernel(orig, stride, width, height, ctu, perRow, out); });
  return 0;
}

// tu_pred_kernel: prediction + residual of TU jobs from the planes
extern "C" int emul_tu_pred(const int16_t* orig, const int16_t* reco, int stride, int bd, int ctu, const vvcb_rmd_visit* visits, const vvcb_tu_src* src,
                            const vvcb_tu_job* jobs, int n, int16_t* pred, int16_t* resi)
{
  static Rom rom;
  fill_rom(rom);
  TuPredParams Q;
  Q.visits = visits; Q.src = src; Q.jobs = jobs; Q.n = n; Q.pred = pred; Q.resi = resi; Q.orig = orig; Q.reco = reco; Q.stride = stride;
  Q.bd = bd; Q.ctu = ctu; Q.rom = &rom;
  emu_launch(2, kTuPredWarps * 32, [&] { tu_pred_kernel(Q); });
  return 0;
}

// rate_kernel on given levels (vvcb_residual_bits)
extern "C" int emul_residual_bits(const vvcb_tu_job* jobs, int n, const int32_t* levels, const vvcb_ctx_states* states, int depQuant, vvcb_tu_result* results)
{
  static DqRom dqRom;
  static RateRom rr;
  fill_dq_rom(dqRom);
  for (int i = 0; i < 512; i++) rr.binFracBits[i] = kBinFracBits[i];
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  RateParams R;
  R.jobs = jobs; R.order = order.data(); R.n = n; R.level = levels; R.results = results; R.states = states; R.rom = &dqRom; R.rate = &rr; R.depQuant = depQuant;
  emu_launch(2, kRateThreads, [&] { rate_kernel(R); });
  return 0;
}

// make_mode_param (vvcb_core.cuh): the per (shape, mode, reference line) prediction parameters the library's ROM is built from
extern "C" void emul_mode_param(int w, int h, int mode, int mrl, int32_t* out)
{
  const ModeParam p = make_mode_param(w, h, mode, mrl);
  out[0] = p.is_ver; out[1] = p.ref_filter; out[2] = p.interp; out[3] = p.pdpc; out[4] = p.angle; out[5] = p.inv_angle; out[6] = p.ang_scale;
}
