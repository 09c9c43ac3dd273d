// vvc_intra_b200 -- texture-measure kernels (sm_100a): both read only the original luma plane.
//
//   ctu_hads_kernel   EncCu::updateCtuDataISlice / xCalcHADs8x8_ISlice (EL/EncCu.cpp:564-675), per CTU as
//                     EncSlice::calCostSliceI walks them (EL/EncSlice.cpp:1276-1298)
//   features_kernel   the 27 classifier inputs of the fork's FAST_ALGORITHM block (EL/EncCu.cpp:72-164, 816-1138)
//
// The reference computes the features with OpenCV on an 8-bit copy of the block (convertTo(CV_8U) saturates);
// the arithmetic below is the integer form of those calls: filter2D = 3x3 correlation with BORDER_REFLECT_101 and
// saturation to [0,255]; `A/4 + B/4 + C/4 + D/4` = three chained addWeighted with round-half-even; meanStdDev =
// population moments (all block sizes are powers of two, so sum/N and sqsum/N are exact in double), and the
// reference squares the returned standard deviation again -- sqrt and square are kept in IEEE double.
#pragma once
#include "vvcb_core.cuh"

namespace {

using namespace vvcb;

constexpr int kFeatThreads = 128;
constexpr int kFeatMaxSide = 64;      // luma CUs of the all-intra dual-tree configuration

// ---------------------------------------------------------------------------------------------------- a16
// One thread per 8x8 block: rows are 16-byte loads, the 2-D Hadamard runs in registers.
__global__ void __launch_bounds__(256) ctu_hads_kernel(const int16_t* __restrict__ orig, int stride, int width, int height, int ctu,
                                                       int ctusPerRow, int32_t* __restrict__ out)
{
  const int log2ctu = vlog2(ctu);
  const int bw = width >> 3, bh = height >> 3;          // 8x8 grid anchored at the picture origin == anchored at each CTU origin
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < bw * bh; b += gridDim.x * blockDim.x) {
    const int bx = b % bw, by = b / bw;
    const int16_t* p = orig + (size_t)(by * 8) * stride + bx * 8;
    int m[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
#if defined(__CUDA_ARCH__)
      const int4 v = *reinterpret_cast<const int4*>(p + (size_t)i * stride);
      const int wv[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for (int k = 0; k < 4; k++) { m[i][2 * k] = (int)(int16_t)(wv[k] & 0xffff); m[i][2 * k + 1] = wv[k] >> 16; }
#else
      for (int k = 0; k < 8; k++) m[i][k] = p[(size_t)i * stride + k];
#endif
    }
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (!(j & s)) { const int a = m[i][j], c = m[i][j + s]; m[i][j] = a + c; m[i][j + s] = a - c; }
    }
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (!(i & s)) { const int a = m[i][j], c = m[i + s][j]; m[i][j] = a + c; m[i + s][j] = a - c; }
    }
    int sum = -vabs(m[0][0]);
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) sum += vabs(m[i][j]);
    // a block counts for its CTU only if it lies completely inside it; CTU sizes are multiples of 8, and the clipped last
    // CTU row/column keeps the blocks with (xBl + 8) <= width (EL/EncCu.cpp:665-667) -- exactly the picture-anchored grid
    atomicAdd(&out[((by * 8) >> log2ctu) * ctusPerRow + ((bx * 8) >> log2ctu)], (sum + 2) >> 2);
  }
}

// ---------------------------------------------------------------------------------------------------- a17
struct FeatParams {
  const vvcb_feat_job* jobs;
  int n;
  vvcb_feat_result* results;
  const int16_t* orig;
  int stride;
};

VHD int sat_u8(int v) { return vmin(vmax(v, 0), 255); }

// saturate_cast<uchar>(cvRound(q4 / 4.0)) for a non-negative numerator in quarter units, round half to even
VHD int round_quarters_u8(int q4)
{
  const int q = q4 >> 2, r = q4 & 3;
  return vmin(q + ((r == 3 || (r == 2 && (q & 1))) ? 1 : 0), 255);
}

// int(stddev * stddev) of cv::meanStdDev from integer moments over n = 2^k samples
VHD int int_var(long long sum, long long sqsum, int n)
{
  const double scale = 1.0 / (double)n;
#if defined(__CUDA_ARCH__)
  const double mean = __dmul_rn((double)sum, scale);
  double v = __dsub_rn(__dmul_rn((double)sqsum, scale), __dmul_rn(mean, mean));
  if (v < 0.0) v = 0.0;
  const double sd = __dsqrt_rn(v);
  return (int)__dmul_rn(sd, sd);
#else
  const double mean = (double)sum * scale;
  double v = (double)sqsum * scale - mean * mean;
  if (v < 0.0) v = 0.0;
  const volatile double sd = sqrt(v);
  return (int)(sd * sd);
#endif
}

VHD int sccd_of(const int* v, int n)
{
  int mean = 0, acc = 0;
  for (int i = 0; i < n; i++) mean += v[i];
  mean /= n;
  for (int i = 0; i < n; i++) acc += (v[i] - mean) * (v[i] - mean);
  return acc / n;
}

// slots of the per-job accumulator array in shared memory
enum { kAccGrad = 0, kAccMadp = 4, kAccMadpSq = 5, kAccMax = 6, kAccNb = 7, kAccCell = 17, kAccCount = 17 + 32 };

__device__ __forceinline__ int warp_sum(int v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kFeatThreads) features_kernel(FeatParams P)
{
  __shared__ uint8_t px[kFeatMaxSide * kFeatMaxSide];
  __shared__ int acc[kAccCount];
  const int lane = threadIdx.x & 31;
  for (int ji = blockIdx.x; ji < P.n; ji += gridDim.x) {
    const vvcb_feat_job job = P.jobs[ji];
    const int w = job.cu.w, h = job.cu.h, n = w * h, lw = vlog2(w), lh = vlog2(h);
    __syncthreads();
    for (int i = threadIdx.x; i < kAccCount; i += kFeatThreads) acc[i] = 0;
    for (int i = threadIdx.x; i < n; i += kFeatThreads) {
      const int x = i & (w - 1), y = i >> lw;
      px[i] = (uint8_t)sat_u8(P.orig[(size_t)(job.cu.y + y) * P.stride + job.cu.x + x]);
    }
    __syncthreads();

    // ---- per-pixel terms: Sobel-like gradients, their rounded mean map, MADP, and the 4x4 grid of partial moments
    int g0 = 0, g1 = 0, g2 = 0, g3 = 0, gmax = 0, ms = 0, mq = 0;
    for (int i = threadIdx.x; i < n; i += kFeatThreads) {
      const int x = i & (w - 1), y = i >> lw;
      const int xm = x == 0 ? 1 : x - 1, xp = x == w - 1 ? w - 2 : x + 1;        // BORDER_REFLECT_101
      const int ym = y == 0 ? 1 : y - 1, yp = y == h - 1 ? h - 2 : y + 1;
      const int a = px[(ym << lw) + xm], b = px[(ym << lw) + x], c = px[(ym << lw) + xp];
      const int d = px[(y << lw) + xm],  e = px[i],              f = px[(y << lw) + xp];
      const int g = px[(yp << lw) + xm], k = px[(yp << lw) + x], l = px[(yp << lw) + xp];
      const int gh   = sat_u8((c - a) + 2 * (f - d) + (l - g));                  // kern_H   (EL/EncCu.cpp:997)
      const int gv   = sat_u8((a - g) + 2 * (b - k) + (c - l));                  // kern_V   (:1000)
      const int g45  = sat_u8((b - d) + 2 * (c - g) + (f - k));                  // kern_45  (:1008)
      const int g135 = sat_u8((b - f) + 2 * (a - l) + (d - k));                  // kern_135 (:1011)
      g0 += gh; g1 += gv; g2 += g45; g3 += g135;
      int t = round_quarters_u8(gh + gv);                                        // :1027, see the header comment
      t = round_quarters_u8(4 * t + g45);
      t = round_quarters_u8(4 * t + g135);
      gmax = vmax(gmax, t);
      // get_madp (:73-134): existing neighbours only (3 / 5 / 8 of them)
      int s = 0, cnt = 0;
      const bool hasL = x > 0, hasR = x < w - 1, hasU = y > 0, hasD = y < h - 1;
      if (hasU) { s += vabs(b - e); cnt++; if (hasL) { s += vabs(a - e); cnt++; } if (hasR) { s += vabs(c - e); cnt++; } }
      if (hasD) { s += vabs(k - e); cnt++; if (hasL) { s += vabs(g - e); cnt++; } if (hasR) { s += vabs(l - e); cnt++; } }
      if (hasL) { s += vabs(d - e); cnt++; }
      if (hasR) { s += vabs(f - e); cnt++; }
      const int m = cnt == 8 ? s >> 3 : (cnt == 5 ? s / 5 : s / 3);
      ms += m; mq += m * m;
      const int cell = (((y << 2) >> lh) << 2) + ((x << 2) >> lw);               // quarter-row x quarter-column
      atomicAdd(&acc[kAccCell + cell], e);
      atomicAdd(&acc[kAccCell + 16 + cell], e * e);
    }
    g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2); g3 = warp_sum(g3); ms = warp_sum(ms); mq = warp_sum(mq);
    for (int o = 16; o > 0; o >>= 1) gmax = vmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0) {
      atomicAdd(&acc[kAccGrad + 0], g0); atomicAdd(&acc[kAccGrad + 1], g1); atomicAdd(&acc[kAccGrad + 2], g2); atomicAdd(&acc[kAccGrad + 3], g3);
      atomicAdd(&acc[kAccMadp], ms); atomicAdd(&acc[kAccMadpSq], mq);
      atomicMax(&acc[kAccMax], gmax);
    }

    // ---- get_context (:137-163): first two moments of every neighbour CU's own area, straight from the plane
    for (int k = 0; k < job.n_neighbours; k++) {
      const vvcb_feat_cu nb = job.nb[k];
      const int nlw = vlog2(nb.w), nn = nb.w * nb.h;
      int s = 0, q = 0;
      for (int i = threadIdx.x; i < nn; i += kFeatThreads) {
        const int v = sat_u8(P.orig[(size_t)(nb.y + (i >> nlw)) * P.stride + nb.x + (i & (nb.w - 1))]);
        s += v; q += v * v;
      }
      s = warp_sum(s); q = warp_sum(q);
      if (lane == 0) { atomicAdd(&acc[kAccNb + 2 * k], s); atomicAdd(&acc[kAccNb + 2 * k + 1], q); }
    }
    __syncthreads();

    if (threadIdx.x == 0) {
      vvcb_feat_result r;
      int* f = r.f;
      for (int i = 0; i < VVCB_NUM_FEATURES; i++) f[i] = 0;
      f[0] = h; f[1] = w; f[2] = job.cu.qt_depth; f[3] = job.cu.mt_depth;
      // G_x = sum / CU_size in double, int() truncates; sums are non-negative and CU_size is a power of two
      f[4] = acc[kAccGrad + 0] >> (lw + lh); f[5] = acc[kAccGrad + 1] >> (lw + lh);
      f[6] = acc[kAccGrad + 2] >> (lw + lh); f[7] = acc[kAccGrad + 3] >> (lw + lh);
      f[8] = (int)(((long long)acc[kAccGrad + 0] + acc[kAccGrad + 1] + acc[kAccGrad + 2] + acc[kAccGrad + 3]) >> (lw + lh + 2));
      f[9] = acc[kAccMax];
      // region moments from the 4x4 grid
      const int* cs = acc + kAccCell;
      const int* cq = acc + kAccCell + 16;
      auto region = [&](int r0, int r1, int c0, int c1) {      // rows [r0,r1) x cols [c0,c1) of the grid
        long long s = 0, q = 0;
        for (int a = r0; a < r1; a++) for (int b = c0; b < c1; b++) { s += cs[a * 4 + b]; q += cq[a * 4 + b]; }
        return int_var(s, q, (n >> 4) * (r1 - r0) * (c1 - c0));
      };
      f[10] = region(0, 4, 0, 4);
      f[11] = int_var(acc[kAccMadp], acc[kAccMadpSq], n);
      if (job.n_neighbours > 0) {
        const int m = job.n_neighbours;
        int vmn = 0x7fffffff, vmx = -1, vs = 0, qmn = 255, qmx = 0, qs = 0, mmn = 255, mmx = 0, msum = 0;
        for (int k = 0; k < m; k++) {
          const int v = int_var(acc[kAccNb + 2 * k], acc[kAccNb + 2 * k + 1], job.nb[k].w * job.nb[k].h);
          vmn = vmin(vmn, v); vmx = vmax(vmx, v); vs += v;
          qmn = vmin(qmn, job.nb[k].qt_depth); qmx = vmax(qmx, job.nb[k].qt_depth); qs += job.nb[k].qt_depth;
          mmn = vmin(mmn, job.nb[k].mt_depth); mmx = vmax(mmx, job.nb[k].mt_depth); msum += job.nb[k].mt_depth;
        }
        f[12] = vmx; f[13] = vmn; f[14] = vs / m;
        f[15] = qmx; f[16] = qmn; f[17] = qs / m;
        f[18] = mmx; f[19] = mmn; f[20] = msum / m;
      }
      int v[4];
      v[0] = region(0, 2, 0, 4); v[1] = region(2, 4, 0, 4);                               f[21] = sccd_of(v, 2);   // BT-H halves
      v[0] = region(0, 4, 0, 2); v[1] = region(0, 4, 2, 4);                               f[22] = sccd_of(v, 2);   // BT-V halves
      v[0] = region(0, 1, 0, 4); v[1] = region(1, 3, 0, 4); v[2] = region(3, 4, 0, 4);    f[23] = sccd_of(v, 3);   // TT-H
      v[0] = region(0, 4, 0, 1); v[1] = region(0, 4, 1, 3); v[2] = region(0, 4, 3, 4);    f[24] = sccd_of(v, 3);   // TT-V
      v[0] = region(0, 2, 0, 2); v[1] = region(0, 2, 2, 4); v[2] = region(2, 4, 0, 2); v[3] = region(2, 4, 2, 4);
      f[25] = sccd_of(v, 4);                                                                                       // QT quadrants
      f[26] = f[10] < f[13] ? 0 : (f[10] > f[12] ? 2 : 1);
      r.valid = job.n_neighbours >= 3;
      P.results[ji] = r;
    }
  }
}

}  // namespace
