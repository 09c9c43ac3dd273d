// Measuring aid (GPU box): reference-line staging of the evaluation kernels, A/B.
//   A  the product's path: build_line_sets<3> (vvcb_rmd.cuh) -- every lane issues plain global loads (LDG, L1 / L2 hits) along the walk
//      of the three reference lines and stores the samples into the warp's shared-memory lines;
//   B  the same walk fed from two shared-memory windows that TMA tensor copies (cp.async.bulk.tensor.2d, one 4-row window above the CU and
//      one 8-column window left of it, both starting on a 16-byte boundary) bring in, double buffered: the copies of visit i+1 are in flight while visit i is assembled.
// Both variants then fold the lines into one checksum per visit (so that nothing is optimised away); the checksums must agree.
// One warp per visit, all candidate CUs of a 1920x1080 picture of every shape the shipped configuration produces, neighbours fully
// available inside the picture -- the staging work of one rough-mode-decision sweep, without the prediction / SAD / SATD that follows it.
//
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/tma_lines_ab.bin tools/tma_lines_ab.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include "../vvc_intra_b200/csrc/vvcb_rmd.cuh"

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kWarps = 4;                       // visits in flight per CTA (the double-buffered windows of B: 6.4 KB per warp)
constexpr int kTopRows = 4, kLeftCols = 8;      // window above: rows y-4 .. y-1; window left: columns x-8 .. x-1
constexpr int kTopMaxW = 2 * 64 + 8, kLeftMaxH = 2 * 64;
// a tensor copy must start on a 16-byte boundary of the plane: the windows start at column (x - 4) & ~7, the walk adds x - that

struct Maps { CUtensorMap top[5], left[5]; };   // by log2(w) - 2 and log2(h) - 2: the box is part of the descriptor

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <class SM> __device__ __forceinline__ int fold_lines(const SM& sm, const Shape& sh, int lane)
{
  int acc = 0;
  for (int q = 0; q < 3; q++) {
    const int set = q ? q + 1 : 0, mrl = q == 2 ? 3 : q;
    for (int i = lane; i < 2 * sh.w + 1 + mrl; i += 32) acc += sm.lines[set][0][i] * (i + 1);
    for (int i = lane; i < 2 * sh.h + 1 + mrl; i += 32) acc += sm.lines[set][1][i] * (i + 3);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

__global__ void __launch_bounds__(kWarps * 32) lines_ldg_kernel(const vvcb_rmd_visit* visits, int n, const int16_t* reco, int stride, int bd, int* out)
{
  __shared__ WarpSmem smem[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = blockIdx.x * kWarps + warp; i < n; i += gridDim.x * kWarps) {
    const vvcb_rmd_visit v = visits[i];
    const Shape sh = make_shape(v.log2w, v.log2h);
    __syncwarp();
    build_line_sets<3>(smem[warp], v, sh, reco, stride, bd, lane);
    __syncwarp();
    const int acc = fold_lines(smem[warp], sh, lane);
    if (lane == 0) out[i] = acc;
  }
}

struct alignas(128) Windows { alignas(128) int16_t top[kTopRows * kTopMaxW]; alignas(128) int16_t left[kLeftMaxH * kLeftCols]; };

__device__ __forceinline__ void tma_issue(const Maps* maps, const vvcb_rmd_visit& v, Windows& win, uint64_t* bar, int mode)
{
  const int lw = (mode & 64) ? 2 : v.log2w, lh = (mode & 64) ? 2 : v.log2h;
  const int cx = (mode & 32) ? 16 : (int)v.x, cy = (mode & 32) ? 16 : (int)v.y;
  const int w = 1 << lw, h = 1 << lh;
  const uint32_t bytes = (uint32_t)((((mode & 1) ? kTopRows * (2 * w + 8) : 0) + ((mode & 2) ? 2 * h * kLeftCols : 0)) * sizeof(int16_t));
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
  if (mode & 1)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(win.top)), "l"(&maps->top[lw - 2]), "r"(b), "r"((cx - 4) & ~7), "r"(cy - kTopRows) : "memory");
  if (mode & 2)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(win.left)), "l"(&maps->left[lh - 2]), "r"(b), "r"((cx - 4) & ~7), "r"(cy) : "memory");
}

__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity)
{
  const uint32_t b = smem_u32(bar);
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(parity) : "memory");
}

// load_line_entry (vvcb_rmd.cuh) with the sample taken from the staged windows
__device__ __forceinline__ int load_line_entry_win(const LineCtx& c, const Windows& win, int topStride, int off, int bd, int i)
{
  const int pos = i < c.g.n ? i : (i < c.g.n + c.extTop ? c.g.n - 1 : 0);
  const int src = line_source(c.g, pos);
  int val = 1 << (bd - 1);
  if (src >= 0) {
    bool isLeft; int k, dx, dy;
    line_pos(c.g, src, isLeft, k, dx, dy);
    val = dy < 0 ? win.top[(dy + kTopRows) * topStride + dx + off] : win.left[dy * kLeftCols + dx + off];
  }
  return val;
}

__global__ void __launch_bounds__(kWarps * 32) lines_tma_kernel(const Maps* __restrict__ maps, const vvcb_rmd_visit* visits, int n, int bd, int* out, int mode)
{
  __shared__ WarpSmem smem[kWarps];
  __shared__ Windows win[kWarps][2];
  __shared__ alignas(8) uint64_t bars[kWarps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    for (int s = 0; s < 2; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[warp][s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int nw = blockDim.x >> 5;
  const int step = gridDim.x * nw;
  int i = blockIdx.x * nw + warp;
  uint32_t phase[2] = { 0, 0 };
  int s = 0;
  const bool ahead = !(mode & 4);
  if (ahead && i < n && lane == 0) tma_issue(maps, visits[i], win[warp][0], &bars[warp][0], mode);
  for (; i < n; i += step, s ^= 1) {
    const vvcb_rmd_visit v = visits[i];
    if (!ahead && lane == 0 && !(mode & 8)) tma_issue(maps, v, win[warp][s], &bars[warp][s], mode);
    if (ahead && i + step < n && lane == 0) tma_issue(maps, visits[i + step], win[warp][s ^ 1], &bars[warp][s ^ 1], mode);     // the next visit's windows fly meanwhile
    const Shape sh = make_shape(v.log2w, v.log2h);
    if (!(mode & 8)) bar_wait(&bars[warp][s], phase[s]);
    phase[s] ^= 1;
    LineCtx c[3];
#pragma unroll
    for (int q = 0; q < 3; q++) c[q] = make_line_ctx(v, sh, q == 2 ? 3 : q);
    const int topStride = 2 * sh.w + 8, off = v.x - ((v.x - 4) & ~7);
    if (!(mode & 16))
    for (int i0 = lane; i0 < c[2].total; i0 += 32)
#pragma unroll
      for (int q = 0; q < 3; q++)
        if (i0 < c[q].total) store_line_entry(smem[warp], q ? q + 1 : 0, c[q], i0, load_line_entry_win(c[q], win[warp][s], topStride, off, bd, i0));
    __syncwarp();
    const int acc = fold_lines(smem[warp], sh, lane);
    if (lane == 0) out[i] = acc;
    __syncwarp();                                   // the window and the lines are free again
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
  const int mode = argc > 1 ? atoi(argv[1]) : 3;
  const int nwB = argc > 2 ? atoi(argv[2]) : kWarps;   // diagnosis: warps per CTA of variant B; mode bit 4: no copy-ahead      // diagnosis: 1 = only the window above, 2 = only the window left

  const int W = 1920, H = 1080, stride = 1920, bd = 10;
  std::vector<int16_t> plane((size_t)stride * H);
  uint32_t r = 12345;
  for (auto& p : plane) { r = r * 1664525u + 1013904223u; p = (int16_t)((r >> 16) & 1023); }
  // every candidate CU of the picture: all aligned positions of the 17 shapes, availability = what lies inside the picture
  std::vector<vvcb_rmd_visit> visits;
  for (int lw = 2; lw <= 6; lw++)
    for (int lh = 2; lh <= 6; lh++) {
      if ((lw == 6) != (lh == 6)) continue;
      const int w = 1 << lw, h = 1 << lh;
      for (int y = 0; y + h <= H; y += h)
        for (int x = 0; x + w <= W; x += w) {
          vvcb_rmd_visit v = {};
          v.x = (int16_t)x; v.y = (int16_t)y; v.log2w = (uint8_t)lw; v.log2h = (uint8_t)lh;
          v.avail_al = x > 0 && y > 0;
          v.n_above = y > 0 ? w / 4 : 0;
          v.n_above_right = y > 0 ? std::max(0, std::min(w, W - x - w)) / 4 : 0;
          v.n_left = x > 0 ? h / 4 : 0;
          v.n_below_left = x > 0 ? std::max(0, std::min(h, H - y - h)) / 4 : 0;
          visits.push_back(v);
        }
    }
  const int n = (int)visits.size();
  int16_t* dPlane; vvcb_rmd_visit* dVisits; int *dOutA, *dOutB;
  CKC(cudaMalloc(&dPlane, plane.size() * 2)); CKC(cudaMalloc(&dVisits, (size_t)n * sizeof(vvcb_rmd_visit)));
  CKC(cudaMalloc(&dOutA, (size_t)n * 4)); CKC(cudaMalloc(&dOutB, (size_t)n * 4));
  CKC(cudaMemcpy(dPlane, plane.data(), plane.size() * 2, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dVisits, visits.data(), (size_t)n * sizeof(vvcb_rmd_visit), cudaMemcpyHostToDevice));

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
  if (!encode) { fprintf(stderr, "cuTensorMapEncodeTiled not available\n"); return 1; }
  Maps maps;
  const cuuint64_t dims[2] = { (cuuint64_t)stride, (cuuint64_t)H }, strides[1] = { (cuuint64_t)stride * 2 };
  const cuuint32_t ones[2] = { 1, 1 };
  for (int k = 0; k < 5; k++) {
    const cuuint32_t boxT[2] = { (cuuint32_t)(2 * (4 << k) + 8), (cuuint32_t)kTopRows }, boxL[2] = { (cuuint32_t)kLeftCols, (cuuint32_t)(2 * (4 << k)) };
    CUresult a = encode(&maps.top[k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dPlane, dims, strides, boxT, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult b = encode(&maps.left[k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dPlane, dims, strides, boxL, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (a != CUDA_SUCCESS || b != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed (%d, %d) for shape class %d\n", (int)a, (int)b, k); return 1; }
  }
  Maps* dMaps; CKC(cudaMalloc(&dMaps, sizeof(Maps))); CKC(cudaMemcpy(dMaps, &maps, sizeof(Maps), cudaMemcpyHostToDevice));     // descriptors in global memory
  cudaDeviceProp prop; CKC(cudaGetDeviceProperties(&prop, 0));
  cudaEvent_t e0, e1; CKC(cudaEventCreate(&e0)); CKC(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto launch) {
    std::vector<float> ms;
    for (int it = 0; it < 12; it++) {
      CKC(cudaEventRecord(e0)); launch(); CKC(cudaEventRecord(e1)); CKC(cudaEventSynchronize(e1));
      float t; CKC(cudaEventElapsedTime(&t, e0, e1)); if (it >= 2) ms.push_back(t);
    }
    CKC(cudaGetLastError());
    std::sort(ms.begin(), ms.end());
    printf("%-46s median %.3f ms  min %.3f ms  (%d visits)\n", name, ms[ms.size() / 2], ms[0], n);
  };
  for (int ctasPerSm : { 4, 6, 8 }) {
    const int grid = prop.multiProcessorCount * ctasPerSm;
    char name[96];
    snprintf(name, sizeof(name), "A  LDG walk (product path), %d CTAs/SM", ctasPerSm);
    timeit(name, [&] { lines_ldg_kernel<<<grid, kWarps * 32>>>(dVisits, n, dPlane, stride, bd, dOutA); });
    snprintf(name, sizeof(name), "B  TMA windows, double buffered, %d CTAs/SM", ctasPerSm);
    timeit(name, [&] { lines_tma_kernel<<<grid, nwB * 32>>>(dMaps, dVisits, n, bd, dOutB, mode); });
  }
  std::vector<int> a(n), b(n);
  CKC(cudaMemcpy(a.data(), dOutA, (size_t)n * 4, cudaMemcpyDeviceToHost)); CKC(cudaMemcpy(b.data(), dOutB, (size_t)n * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < n; i++) bad += a[i] != b[i];
  printf("checksums: %d of %d visits differ\n", bad, n);
  return bad != 0;
}
