// vvc_intra_b200 -- RMD device kernels (sm_100a).  Included by vvcb_api.cu (the product) and, through a
// CUDA-on-pthreads shim, by tests/host_emul/emul_rmd.cpp (test-only lane emulation of this very source).
//
// Pipeline of one vvcb_rmd_eval call (all on the context's stream):
//   rmd_plan_kernel    one thread per visit: cuts the visit into work items of <= kItemTasks lane-tasks
//   rmd_eval_kernel    persistent warps pull work items; a lane predicts one 8x8 (or 4x4) unit of one
//                      evaluation slot, keeps the residual in registers, computes SAD and the Walsh-
//                      Hadamard SATD there, and the slot's lanes reduce with warp shuffles
//   rmd_lists_kernel   one thread per visit: mode bits, double-precision costs and the exact replay of
//                      the reference's candidate-list insertions
#pragma once
#include "vvcb_core.cuh"

using namespace vvcb;

// =====================================================================================================
// device helpers
// =====================================================================================================
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads     = kWarpsPerCta * 32;

struct WarpSmem {
  int16_t lines[kNumSets][2][kLineMax];   // [set][0 top / 1 left][index]
  int16_t slotLines[kSlotLineWords];      // per-slot main lines / MIP reduced predictions in flight
  int     mipBnd[8];                      // Haar-averaged boundary: [0..4) top, [4..8) left
};

struct EvalParams {
  const vvcb_rmd_visit* visits;
  const WorkItem*       items;
  const unsigned*       itemCount;
  unsigned*             cursor;
  vvcb_rmd_detail*      details;  // sad/satd tables, one per visit
  const int16_t*        orig;
  const int16_t*        reco;
  int                   stride;   // both planes
  int                   bd, ctu;
  const Rom*            rom;
  int16_t*              predOut;  // optional: [active slot][h][w] of visit 0 (debug / parity)
};

__device__ __forceinline__ int active_slot_to_slot(int a, bool mrlAllowed)
{
  if (a < VVCB_NUM_LUMA_MODE) return a;
  a -= VVCB_NUM_LUMA_MODE;
  if (mrlAllowed) { if (a < 10) return VVCB_SLOT_MRL1 + a; a -= 10; }
  return VVCB_SLOT_MIP + a;
}

__device__ __forceinline__ int num_active_slots(bool mrlAllowed, int numMip)
{
  return VVCB_NUM_LUMA_MODE + (mrlAllowed ? 10 : 0) + numMip;
}

__device__ __forceinline__ bool visit_mrl_allowed(const vvcb_rmd_visit& v, int ctu)
{
  return !(v.flags & VVCB_VISIT_NO_MRL) && (v.y & (ctu - 1)) != 0;
}

__device__ __forceinline__ int visit_num_mip(const vvcb_rmd_visit& v)
{
  return (v.flags & VVCB_VISIT_NO_MIP) ? 0 : mip_num_modes(1 << v.log2w, 1 << v.log2h);
}

// ---- reference lines of one visit into the warp's shared memory ---------------------------------------
__device__ void build_line_set(WarpSmem& sm, int set, int mrl, const vvcb_rmd_visit& v, const Shape& sh,
                               const int16_t* reco, int stride, int bd, int lane)
{
  const LineGeom g = make_line_geom(v, sh.w, sh.h, mrl);
  const int16_t* base = reco + (size_t)v.y * stride + v.x;
  for (int i = lane; i < g.n; i += 32) {
    const int src = line_source(g, i);
    int val = 1 << (bd - 1);
    bool isLeft; int k, dx, dy;
    if (src >= 0) {
      line_pos(g, src, isLeft, k, dx, dy);
      val = base[dy * stride + dx];
    }
    line_pos(g, i, isLeft, k, dx, dy);
    if (isLeft) sm.lines[set][1][k] = (int16_t)val;
    else {
      sm.lines[set][0][k] = (int16_t)val;
      if (k == 0) sm.lines[set][1][0] = (int16_t)val;
    }
  }
}

__device__ void build_filtered_set(WarpSmem& sm, const Shape& sh, int lane)
{
  const int nTop = 2 * sh.w + 1, nLeft = 2 * sh.h + 1;
  const int16_t* top = sm.lines[0][0];
  const int16_t* left = sm.lines[0][1];
  for (int i = lane; i < nTop; i += 32) {
    int v;
    if (i == 0) v = (left[1] + 2 * top[0] + top[1] + 2) >> 2;
    else if (i == nTop - 1) v = top[i];
    else v = (top[i - 1] + 2 * top[i] + top[i + 1] + 2) >> 2;
    sm.lines[1][0][i] = (int16_t)v;
    if (i == 0) sm.lines[1][1][0] = (int16_t)v;
  }
  for (int i = 1 + lane; i < nLeft; i += 32) {
    int v;
    if (i == nLeft - 1) v = left[i];
    else v = (left[i - 1] + 2 * left[i] + left[i + 1] + 2) >> 2;   // left[0] == top[0]
    sm.lines[1][1][i] = (int16_t)v;
  }
}

__device__ void build_mip_boundary(WarpSmem& sm, const Shape& sh, const MipGeom& mg, int lane)
{
  if (lane < 8) {
    const int side = lane >> 2, i = lane & 3;          // 0 top, 1 left
    if (i < mg.bsz) {
      const int len = side ? sh.h : sh.w;
      const int f = len / mg.bsz;
      const int16_t* src = sm.lines[0][side] + 1 + i * f;
      int s = 0;
      for (int k = 0; k < f; k++) s += src[k];
      sm.mipBnd[side * 4 + i] = f == 1 ? s : (s + (f >> 1)) >> vlog2(f);
    }
  }
}

// ---- per-slot set-up by the slot's lane group -----------------------------------------------------------
__device__ __forceinline__ SlotInfo make_slot_info(const Rom& rom, const vvcb_rmd_visit& v, const Shape& sh, int slot)
{
  SlotInfo s;
  s.mrl = 0; s.set = 0;
  if (slot >= VVCB_SLOT_MIP) { s.kind = 3; s.mode = slot - VVCB_SLOT_MIP; s.p = ModeParam{}; return s; }
  if (slot >= VVCB_SLOT_MRL1) {
    const int li = slot >= VVCB_SLOT_MRL3 ? 1 : 0;
    s.mrl  = li ? 3 : 1;
    s.set  = 2 + li;
    s.mode = v.mpm[1 + (slot - VVCB_SLOT_MRL1) % 5];
    s.p    = rom.mode[sh.lw - 2][sh.lh - 2][s.mode];
    s.p.ref_filter = 0; s.p.interp = 0; s.p.pdpc = 0;      // CL/IntraPrediction.cpp:509, :559-575
  } else {
    s.mode = slot;
    s.p    = rom.mode[sh.lw - 2][sh.lh - 2][slot];
    s.set  = s.p.ref_filter ? 1 : 0;
  }
  s.kind = s.mode == 0 ? 0 : (s.mode == 1 ? 1 : 2);
  return s;
}

// main line of an angular slot: ml[t + off] = refMain0[t]  (CL/IntraPrediction.cpp:654-726)
__device__ void build_slot_line(int16_t* ml, const WarpSmem& sm, const SlotInfo& s, const Shape& sh, int gl, int gsize)
{
  const int16_t* mainSrc = sm.lines[s.set][s.p.is_ver ? 0 : 1];
  const int16_t* sideSrc = sm.lines[s.set][s.p.is_ver ? 1 : 0];
  const int mw = s.p.is_ver ? sh.w : sh.h, mh = s.p.is_ver ? sh.h : sh.w;
  const int mrl = s.mrl;
  if (s.p.angle < 0) {
    const int n = mh + mw + 2 + mrl;                 // t in [-mh, mw + 1 + mrl]
    const int inv = s.p.inv_angle;
    for (int i = gl; i < n; i += gsize) {
      const int t = i - mh;
      ml[i] = t >= 0 ? mainSrc[t] : sideSrc[vmin((-t * inv + 256) >> 9, mh)];
    }
  } else {
    const int mainLen = 2 * mw + mrl;
    const int sft = vmax(0, vlog2(mw) - vlog2(mh));
    const int n = mainLen + 1 + (mrl << sft) + 2;
    for (int i = gl; i < n; i += gsize) ml[i] = mainSrc[vmin(i, mainLen)];
  }
}

__device__ void build_mip_reduced(int16_t* red, const Rom& rom, const WarpSmem& sm, const MipGeom& mg, const Shape& sh,
                                  int bd, int mode, int gl, int gsize)
{
  const int n = mg.redW * mg.redH;
  for (int i = gl; i < n; i += gsize)
    red[i] = (int16_t)mip_reduced_sample(rom, mg, sm.mipBnd, sh.w, sh.h, bd, mode, i % mg.redW, i / mg.redW);
}

// ---- one lane, one unit: prediction in block orientation ---------------------------------------------------
template <int S>
__device__ __forceinline__ void predict_unit(const WarpSmem& sm, const int16_t* slotLine, const SlotInfo& s, const Shape& sh,
                                             const MipGeom& mg, const uint32_t* filt, int bd, int x0, int y0, int (&b)[S][S])
{
  const int maxv = (1 << bd) - 1;
  if (s.kind == 2) {
    const bool ver = s.p.is_ver;
    const int mw = ver ? sh.w : sh.h, mh = ver ? sh.h : sh.w;
    const int off = s.p.angle < 0 ? mh : 0;
    int q[S][S];
    pred_angular_unit<S>(slotLine + off, sm.lines[s.set][ver ? 1 : 0], s.p, s.mrl, mw, mh,
                         ver ? x0 : y0, ver ? y0 : x0, filt, maxv, q);
#pragma unroll
    for (int i = 0; i < S; i++)
#pragma unroll
      for (int j = 0; j < S; j++) b[i][j] = ver ? q[i][j] : q[j][i];
  } else if (s.kind < 2) {
    const int16_t* top = sm.lines[s.set][0];
    const int16_t* left = sm.lines[s.set][1];
    int dc = 0;
    if (s.kind == 1) {                                // CL/IntraPrediction.cpp:248-285
      int sum = 0;
      const int denom = sh.w == sh.h ? 2 * sh.w : vmax(sh.w, sh.h);
      if (sh.w >= sh.h) for (int i = 0; i < sh.w; i++) sum += top[s.mrl + 1 + i];
      if (sh.w <= sh.h) for (int i = 0; i < sh.h; i++) sum += left[s.mrl + 1 + i];
      dc = (sum + (denom >> 1)) >> vlog2(denom);
    }
    pred_planar_dc_unit<S>(top, left, s.kind, s.p.pdpc, dc, sh.lw, sh.lh, x0, y0, b);
  } else {
    const int16_t* top = sm.lines[0][0];
    const int16_t* left = sm.lines[0][1];
    const bool up = mg.upH > 1 || mg.upV > 1;
#pragma unroll
    for (int i = 0; i < S; i++)
#pragma unroll
      for (int j = 0; j < S; j++)
        b[i][j] = up ? mip_upsampled_sample(mg, slotLine, top, left, sh.w, sh.h, x0 + j, y0 + i)
                     : slotLine[(y0 + i) * mg.redW + x0 + j];
  }
}

template <int S>
__device__ __forceinline__ void residual_unit(const int16_t* org, int stride, const int (&b)[S][S], int (&d)[S][S], int& sad)
{
#pragma unroll
  for (int i = 0; i < S; i++) {
    const int16_t* row = org + i * stride;
#pragma unroll
    for (int j = 0; j < S; j += 4) {
      const uint2 v = *reinterpret_cast<const uint2*>(row + j);       // CU positions are multiples of 4 samples
      d[i][j + 0] = (int)(int16_t)(v.x & 0xffff) - b[i][j + 0];
      d[i][j + 1] = (int)(int16_t)(v.x >> 16)    - b[i][j + 1];
      d[i][j + 2] = (int)(int16_t)(v.y & 0xffff) - b[i][j + 2];
      d[i][j + 3] = (int)(int16_t)(v.y >> 16)    - b[i][j + 3];
    }
  }
#pragma unroll
  for (int i = 0; i < S; i++)
#pragma unroll
    for (int j = 0; j < S; j++) sad += vabs(d[i][j]);
}

template <int S>
__device__ __forceinline__ void store_pred(int16_t* out, int w, int x0, int y0, const int (&b)[S][S])
{
#pragma unroll
  for (int i = 0; i < S; i++)
#pragma unroll
    for (int j = 0; j < S; j++) out[(y0 + i) * w + x0 + j] = (int16_t)b[i][j];
}

// =====================================================================================================
// kernels
// =====================================================================================================
__global__ void rmd_plan_kernel(const vvcb_rmd_visit* visits, int n, int ctu, WorkItem* items, unsigned* itemCount)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const vvcb_rmd_visit v = visits[i];
  const Shape sh = make_shape(v.log2w, v.log2h);
  const int nAct = num_active_slots(visit_mrl_allowed(v, ctu), visit_num_mip(v));
  const int perItem = sh.lanes >= kItemTasks ? 1 : kItemTasks / sh.lanes;
  const int nItems = (nAct + perItem - 1) / perItem;
  const unsigned base = atomicAdd(itemCount, (unsigned)nItems);
  for (int k = 0; k < nItems; k++) {
    WorkItem w;
    w.visit = (uint32_t)i;
    w.slot_begin = (uint16_t)(k * perItem);
    w.slot_count = (uint16_t)vmin(perItem, nAct - k * perItem);
    items[base + k] = w;
  }
}

__global__ void __launch_bounds__(kThreads, 2) rmd_eval_kernel(EvalParams P)
{
  __shared__ WarpSmem smem[kWarpsPerCta];
  __shared__ uint32_t sFilt[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpSmem& sm = smem[warp];
  if (threadIdx.x < 64) sFilt[threadIdx.x] = (&P.rom->filt[0][0])[threadIdx.x];
  __syncthreads();
  const Rom& rom = *P.rom;
  const unsigned nItems = *P.itemCount;

  for (;;) {
    unsigned it = 0;
    if (lane == 0) it = atomicAdd(P.cursor, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= nItems) break;
    const WorkItem item = P.items[it];
    const vvcb_rmd_visit v = P.visits[item.visit];
    const Shape sh = make_shape(v.log2w, v.log2h);
    const MipGeom mg = make_mip_geom(sh.w, sh.h);
    const bool mrlAllowed = visit_mrl_allowed(v, P.ctu);
    uint32_t* sadOut  = P.details[item.visit].sad;
    uint32_t* satdOut = P.details[item.visit].satd;
    const int16_t* org = P.orig + (size_t)v.y * P.stride + v.x;

    // ---- reference lines needed by this item's slots
    const int firstSlot = active_slot_to_slot(item.slot_begin, mrlAllowed);
    const int lastSlot  = active_slot_to_slot(item.slot_begin + item.slot_count - 1, mrlAllowed);
    __syncwarp();
    build_line_set(sm, 0, 0, v, sh, P.reco, P.stride, P.bd, lane);
    if (firstSlot < VVCB_SLOT_MRL3 && lastSlot >= VVCB_SLOT_MRL1) build_line_set(sm, 2, 1, v, sh, P.reco, P.stride, P.bd, lane);
    if (firstSlot < VVCB_SLOT_MIP && lastSlot >= VVCB_SLOT_MRL3)  build_line_set(sm, 3, 3, v, sh, P.reco, P.stride, P.bd, lane);
    __syncwarp();
    if (firstSlot < VVCB_SLOT_MRL1) build_filtered_set(sm, sh, lane);
    if (lastSlot >= VVCB_SLOT_MIP)  build_mip_boundary(sm, sh, mg, lane);
    __syncwarp();

    const int lanes = sh.lanes;                       // lanes per slot
    const int gsize = lanes > 32 ? 32 : lanes;        // lanes of one slot inside this warp iteration
    const int gidx  = lane / gsize, gl = lane % gsize;
    const int lineStride = kSlotLineWords / (32 / gsize);
    int16_t* slotLine = sm.slotLines + gidx * lineStride;
    const int nTasks = item.slot_count * lanes;
    int accSad = 0, accSatd = 0;

    for (int base = 0; base < nTasks; base += 32) {
      const int task = base + lane;
      const bool act = task < nTasks;
      const int a = item.slot_begin + (act ? task : nTasks - 1) / lanes;
      const int u = (act ? task : nTasks - 1) % lanes;
      const int slot = active_slot_to_slot(a, mrlAllowed);
      const SlotInfo s = make_slot_info(rom, v, sh, slot);

      __syncwarp();
      if (s.kind == 2) build_slot_line(slotLine, sm, s, sh, gl, gsize);
      else if (s.kind == 3) build_mip_reduced(slotLine, rom, sm, mg, sh, P.bd, s.mode, gl, gsize);
      __syncwarp();

      int sad = 0, satd = 0;
      if (sh.S == 8) {
        const int ux = u % sh.unitsX, uy = u / sh.unitsX;
        const int x0 = ux * 8, y0 = uy * 8;
        int b[8][8], d[8][8];
        predict_unit<8>(sm, slotLine, s, sh, mg, sFilt, P.bd, x0, y0, b);
        if (P.predOut && act) store_pred<8>(P.predOut + (size_t)a * sh.w * sh.h, sh.w, x0, y0, b);
        residual_unit<8>(org + y0 * P.stride + x0, P.stride, b, d, sad);
        wht_rows<8>(d);
        if (sh.tile == 3) {
          satd = (wht_cols_abs_sum<8>(d) + 2) >> 2;                       // CL/RdCost.cpp:2306
        } else {
          // 16x8 / 8x16: the partner lane holds the other 8x8 half; the last butterfly stage across the
          // halves is folded into the absolute sum: |a+b| + |a-b| = 2 max(|a|, |b|)
          wht_cols<8>(d);
          const int pm = sh.tile == 4 ? 1 : sh.unitsX;
          int t = 0;
#pragma unroll
          for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const int mine = vabs(d[i][j]);
              const int other = __shfl_xor_sync(0xffffffffu, mine, pm);
              t += vmax(mine, other);
            }
          const bool owner = sh.tile == 4 ? (ux & 1) == 0 : (uy & 1) == 0;
          satd = owner ? satd_norm_rect(2 * t, true) : 0;                 // CL/RdCost.cpp:2452, :2589
        }
      } else {
        // 4xN / Nx4 shapes: a lane owns one SATD tile = one (4x4) or two (8x4, 4x8) 4x4 units
        const int tilesX = sh.tile == 1 ? sh.w / 8 : sh.w / 4;
        const int tx = u % tilesX, ty = u / tilesX;
        const int tw = sh.tile == 1 ? 8 : 4, th = sh.tile == 2 ? 8 : 4;
        const int nUnits = sh.tile == 0 ? 1 : 2;
        int c0[4][4];
        int t = 0;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (k < nUnits) {
            const int x0 = tx * tw + (sh.tile == 1 ? 4 * k : 0);
            const int y0 = ty * th + (sh.tile == 2 ? 4 * k : 0);
            int b[4][4], d[4][4];
            predict_unit<4>(sm, slotLine, s, sh, mg, sFilt, P.bd, x0, y0, b);
            if (P.predOut && act) store_pred<4>(P.predOut + (size_t)a * sh.w * sh.h, sh.w, x0, y0, b);
            residual_unit<4>(org + y0 * P.stride + x0, P.stride, b, d, sad);
            wht_rows<4>(d);
            if (sh.tile == 0) {
              satd = (wht_cols_abs_sum<4>(d) + 1) >> 1;                     // CL/RdCost.cpp:2209
            } else {
              wht_cols<4>(d);
              if (k == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                  for (int j = 0; j < 4; j++) c0[i][j] = vabs(d[i][j]);
              } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                  for (int j = 0; j < 4; j++) t += vmax(c0[i][j], vabs(d[i][j]));
              }
            }
          }
        }
        if (sh.tile != 0) satd = satd_norm_rect(2 * t, false);             // CL/RdCost.cpp:2662, :2741
      }
      if (!act) { sad = 0; satd = 0; }

      if (lanes <= 32) {
        for (int o = gsize >> 1; o > 0; o >>= 1) {
          sad  += __shfl_xor_sync(0xffffffffu, sad, o);
          satd += __shfl_xor_sync(0xffffffffu, satd, o);
        }
        if (act && gl == 0) { sadOut[slot] = (uint32_t)sad; satdOut[slot] = (uint32_t)satd; }
      } else {
        // 64 lanes per slot: two consecutive warp iterations belong to the same slot
        accSad += sad; accSatd += satd;
        if (((base >> 5) & 1) == 1) {
          for (int o = 16; o > 0; o >>= 1) {
            accSad  += __shfl_xor_sync(0xffffffffu, accSad, o);
            accSatd += __shfl_xor_sync(0xffffffffu, accSatd, o);
          }
          if (lane == 0) { sadOut[slot] = (uint32_t)accSad; satdOut[slot] = (uint32_t)accSatd; }
          accSad = 0; accSatd = 0;
        }
      }
    }
  }
}

// g_aucIntraModeNumFast_UseMPM_2D, CL/Rom.cpp:536
__constant__ uint8_t cFastModes[6][6] = {
  { 3, 3, 3, 3, 2, 2 }, { 3, 3, 3, 3, 3, 2 }, { 3, 3, 3, 3, 3, 2 }, { 3, 3, 3, 3, 3, 2 }, { 2, 3, 3, 3, 3, 2 }, { 2, 2, 2, 2, 2, 3 } };

__device__ void store_list(const CandList& L, int32_t* n, vvcb_mode* m, double* c, int cap)
{
  *n = L.n;
  for (int i = 0; i < cap; i++) {
    if (i < L.n) { m[i] = L.m[i]; if (c) c[i] = L.c[i]; }
    else         { m[i] = mk_mode(0, 0, 0); if (c) c[i] = 0.0; }
  }
}

// One thread per visit: EL/IntraSearch.cpp:489-802 minus the predictions (already reduced to SAD/SATD).
__global__ void rmd_lists_kernel(const vvcb_rmd_visit* visits, int n, int ctu, vvcb_rmd_result* results, vvcb_rmd_detail* details)
{
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= n) return;
  const vvcb_rmd_visit v = visits[vi];
  vvcb_rmd_result& R = results[vi];
  vvcb_rmd_detail& D = details[vi];
  R.pad = 0;
  const int w = 1 << v.log2w, h = 1 << v.log2h;
  const bool mipEnabled = !(v.flags & VVCB_VISIT_NO_MIP);
  const int numMip = visit_num_mip(v);
  const bool testMip = numMip > 0;
  const bool mrlAllowed = visit_mrl_allowed(v, ctu);

  auto dist_of = [&](int slot) -> double {
    const uint64_t sad = D.sad[slot], satd = D.satd[slot];
    return (double)(sad * 2 < satd ? sad * 2 : satd);                        // :515
  };
  auto cost_of = [&](int slot, bool isMip, int mrl, int mode) -> double {
    const uint64_t bits = mode_bits(v.rates, v.mpm, w, h, mrlAllowed, mipEnabled, isMip, mrl, mode);
    return __dadd_rn(dist_of(slot), __dmul_rn((double)bits, v.sqrt_lambda)); // :526, no FMA contraction
  };

  // slots that were not evaluated
  if (!mrlAllowed) for (int s = VVCB_SLOT_MRL1; s < VVCB_SLOT_MIP; s++) { D.sad[s] = VVCB_SAT_NONE; D.satd[s] = VVCB_SAT_NONE; }
  for (int s = VVCB_SLOT_MIP + numMip; s < VVCB_NUM_SLOTS; s++) { D.sad[s] = VVCB_SAT_NONE; D.satd[s] = VVCB_SAT_NONE; }

  int K = cFastModes[v.log2w - 2][v.log2h - 2];
  if (testMip) K += vmax(K, vlog2(vmin(w, h)) - 1);                          // :472
  const int numHad = testMip ? 6 : 3;
  CandList rd, had;
  rd.n = 0; had.n = 0;
  uint64_t checked0 = 0, checked1 = 0;                                       // bSatdChecked
  for (int m = 0; m < VVCB_NUM_LUMA_MODE; m++) {                             // :489-532
    if (m > 1 && (m & 1)) continue;
    if (m < 64) checked0 |= 1ull << m; else checked1 |= 1ull << (m - 64);
    cand_push(rd, mk_mode(0, 0, m), cost_of(m, false, 0, m), K);
    cand_push(had, mk_mode(0, 0, m), dist_of(m), numHad);
  }
  uint8_t parent[VVCB_MAX_LIST];
  for (int i = 0; i < K; i++) parent[i] = rd.m[i].mode;
  for (int i = 0; i < K; i++) {                                              // :577-623
    const int pm = parent[i];
    if (pm > 2 && pm < 66)
      for (int dlt = -1; dlt <= 1; dlt += 2) {
        const int m = pm + dlt;
        const bool done = m < 64 ? (checked0 >> m) & 1 : (checked1 >> (m - 64)) & 1;
        if (done) continue;
        cand_push(rd, mk_mode(0, 0, m), cost_of(m, false, 0, m), K);
        cand_push(had, mk_mode(0, 0, m), dist_of(m), numHad);
        if (m < 64) checked0 |= 1ull << m; else checked1 |= 1ull << (m - 64);
      }
  }
  if (mrlAllowed)                                                            // :635-681
    for (int li = 0; li < 2; li++)
      for (int i = 1; i < 6; i++) {
        const int slot = (li ? VVCB_SLOT_MRL3 : VVCB_SLOT_MRL1) + i - 1;
        const int mrl = li ? 3 : 1;
        cand_push(rd, mk_mode(0, mrl, v.mpm[i]), cost_of(slot, false, mrl, v.mpm[i]), K);
        cand_push(had, mk_mode(0, mrl, v.mpm[i]), dist_of(slot), numHad);
      }
  store_list(rd, &D.n_reg, D.reg_mode, D.reg_cost, VVCB_MAX_LIST);
  store_list(had, &D.n_reg_had, D.reg_had_mode, D.reg_had_cost, VVCB_MAX_HAD_LIST);

  if (testMip) {                                                             // :704-751
    double c3[6];                                                            // costs of MIP modes 3,4,5 and their transposes
    const int off = numMip / 2;
    for (int m = 0; m < numMip; m++) {
      const int slot = VVCB_SLOT_MIP + m;
      const double c = cost_of(slot, true, 0, m);
      if (m >= 3 && m <= 5) c3[m - 3] = c;
      if (m >= 3 + off && m <= 5 + off) c3[3 + m - 3 - off] = c;
      cand_push(rd, mk_mode(1, 0, m), c, K + 1);
      cand_push(had, mk_mode(1, 0, m), __dmul_rn(0.8, dist_of(slot)), numHad);
    }
    // reduceHadCandList, :4333-4405
    const double thr = __dadd_rn(1.0, __ddiv_rn(1.4, __dsqrt_rn((double)(w * h))));
    const int maxPerType = K >> 1;
    const double minCost = rd.c[0];
    CandList tmp;
    tmp.n = 0;
    bool keepOne = rd.n > K;
    int numConv = 0, numMipKept = 0;
    for (int idx = 0; idx < rd.n - (keepOne ? 0 : 1); idx++) {
      bool add;
      if (!rd.m[idx].mip) { add = numConv < 3; numConv += add; }
      else {
        add = numMipKept < maxPerType || rd.c[idx] < __dmul_rn(thr, minCost) || keepOne;
        keepOne = false;
        numMipKept += add;
      }
      if (add) { tmp.m[tmp.n] = rd.m[idx]; tmp.c[tmp.n] = rd.c[idx]; tmp.n++; }
    }
    if (w > 8 && h > 8) {
      CandList srt;
      srt.n = 0;
      for (int m = 3; m <= 5; m++) {
        const bool tr = c3[3 + m - 3] < c3[m - 3];
        cand_push(srt, mk_mode(1, 0, tr ? m + off : m), tr ? c3[3 + m - 3] : c3[m - 3], 3);
      }
      const int baseN = tmp.n;
      for (int idx = 0; idx < 3; idx++) {
        bool inc = false;
        for (int i = 0; i < baseN; i++) inc = inc || same_mode(tmp.m[i], srt.m[idx]);
        if (!inc) { tmp.m[tmp.n] = srt.m[idx]; tmp.c[tmp.n] = 0.0; tmp.n++; break; }   // FastMIP: first one only
      }
    }
    rd = tmp;
    K = rd.n;
  }
  store_list(rd, &R.n_rd, R.rd_mode, R.rd_cost, VVCB_MAX_LIST);
  store_list(had, &R.n_had, R.had_mode, R.had_cost, VVCB_MAX_HAD_LIST);

  for (int i = 0; i < v.num_mpm_cand; i++) {                                 // :777-802
    const vvcb_mode mp = mk_mode(0, 0, v.mpm[i]);
    bool inc = false;
    for (int j = 0; j < K; j++) inc = inc || same_mode(mp, rd.m[j]);
    if (!inc) { rd.m[rd.n] = mp; rd.c[rd.n] = 0.0; rd.n++; K++; }
  }
  store_list(rd, &R.n_final, R.final_mode, nullptr, VVCB_MAX_LIST);
}

}  // namespace
