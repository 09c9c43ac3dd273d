// Measuring aid (GPU box): how long after a kernel of known length the host learns that it is done -- spinning stream synchronize,
// blocking event synchronize, and an event-query loop that sleeps between polls.  nvcc -O2 -o /tmp/sync_latency tools/sync_latency.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <time.h>
#include <algorithm>
#include <vector>
__global__ void spin_kernel(long long cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
static double now_us() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
int main()
{
  cudaStream_t s; cudaStreamCreate(&s);
  cudaEvent_t eb, en, t0, t1; cudaEventCreateWithFlags(&eb, cudaEventBlockingSync | cudaEventDisableTiming); cudaEventCreateWithFlags(&en, cudaEventDisableTiming);
  cudaEventCreate(&t0); cudaEventCreate(&t1);
  for (long long cyc : { 100000ll, 400000ll, 1000000ll }) {
    for (int mode = 0; mode < 4; mode++) {
      std::vector<double> host, dev;
      for (int it = 0; it < 60; it++) {
        const double a = now_us();
        cudaEventRecord(t0, s);
        spin_kernel<<<1, 32, 0, s>>>(cyc);
        cudaEventRecord(t1, s);
        if (mode == 0) cudaStreamSynchronize(s);
        else if (mode == 1) { cudaEventRecord(eb, s); cudaEventSynchronize(eb); }
        else {
          cudaEventRecord(en, s);
          timespec nap = { 0, mode == 2 ? 20000 : 50000 };
          while (cudaEventQuery(en) == cudaErrorNotReady) nanosleep(&nap, nullptr);
        }
        const double b = now_us();
        float ms = 0; cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1);
        if (it >= 10) { host.push_back(b - a); dev.push_back(ms * 1e3); }
      }
      std::sort(host.begin(), host.end()); std::sort(dev.begin(), dev.end());
      printf("kernel %7lld cycles  %-28s host p50 %7.1f us  p90 %7.1f us   device p50 %7.1f us\n", cyc,
             mode == 0 ? "stream synchronize (spin)" : mode == 1 ? "blocking event synchronize" : mode == 2 ? "query + nanosleep 20 us" : "query + nanosleep 50 us",
             host[host.size() / 2], host[host.size() * 9 / 10], dev[dev.size() / 2]);
    }
  }
  return 0;
}
