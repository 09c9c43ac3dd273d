// TEST INFRASTRUCTURE -- a minimal CUDA-on-pthreads shim: one host thread per CUDA thread, warp and
// block barriers from pthread barriers, warp shuffles through a per-warp exchange buffer.  It exists
// so that the *identical* kernel source (vvc_intra_b200/csrc/vvcb_rmd.cuh) can be executed and debugged
// in a container without a GPU.  It is never part of the product library.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <math.h>
#include <atomic>
#include <vector>
#include <functional>

struct emu_dim3 { unsigned x, y, z; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };

struct EmuWarp { pthread_barrier_t bar; int xbuf[32]; };
struct EmuBlock { pthread_barrier_t bar; std::vector<EmuWarp> warps; };

extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local EmuBlock* emuBlock;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __constant__ static const
#define __launch_bounds__(...)

static inline void __syncthreads() { pthread_barrier_wait(&emuBlock->bar); }
static inline EmuWarp& emu_warp() { return emuBlock->warps[threadIdx.x >> 5]; }
static inline void __syncwarp() { pthread_barrier_wait(&emu_warp().bar); }
static inline int __shfl_sync(unsigned, int v, int src)
{
  EmuWarp& w = emu_warp();
  w.xbuf[threadIdx.x & 31] = v;
  pthread_barrier_wait(&w.bar);
  const int r = w.xbuf[src & 31];
  pthread_barrier_wait(&w.bar);
  return r;
}
static inline unsigned __shfl_sync(unsigned m, unsigned v, int src) { return (unsigned)__shfl_sync(m, (int)v, src); }
static inline int __shfl_xor_sync(unsigned, int v, int mask)
{
  EmuWarp& w = emu_warp();
  const int lane = threadIdx.x & 31;
  w.xbuf[lane] = v;
  pthread_barrier_wait(&w.bar);
  const int r = w.xbuf[(lane ^ mask) & 31];
  pthread_barrier_wait(&w.bar);
  return r;
}
static inline unsigned __ballot_sync(unsigned, bool pred)
{
  EmuWarp& w = emu_warp();
  w.xbuf[threadIdx.x & 31] = pred;
  pthread_barrier_wait(&w.bar);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(w.xbuf[i] != 0) << i;
  pthread_barrier_wait(&w.bar);
  return r;
}
static inline int __reduce_max_sync(unsigned, int v)
{
  EmuWarp& w = emu_warp();
  w.xbuf[threadIdx.x & 31] = v;
  pthread_barrier_wait(&w.bar);
  int r = w.xbuf[0];
  for (int i = 1; i < 32; i++) r = w.xbuf[i] > r ? w.xbuf[i] : r;
  pthread_barrier_wait(&w.bar);
  return r;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int* p, int v)
{
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline long long __double_as_longlong(double a) { long long r; __builtin_memcpy(&r, &a, 8); return r; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }

// runs fn() once per thread of a grid x block launch, block after block
void emu_launch(unsigned grid, unsigned block, const std::function<void()>& fn);
