// sync_probe.cu -- how a host learns that a small kernel has finished: cudaStreamSynchronize against spinning on a flag the
// kernel writes into mapped pinned host memory (after a system-scope fence).  Host clock, p50 over 2000 launches each.
// nvcc -O2 -o sync_probe sync_probe.cu && ./sync_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <algorithm>
#include <vector>
static double now_us() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
__global__ void k_work(float* x, volatile unsigned* flag, unsigned seq, int spin) {
  float v = x[threadIdx.x];
  for (int i = 0; i < spin; i++) v = v * 1.0001f + 0.5f;
  x[threadIdx.x] = v;
  __syncthreads();
  if (flag && threadIdx.x == 0) {
    __threadfence_system();
    *flag = seq;
  }
}
int main() {
  float* x;
  cudaMalloc(&x, 1024 * sizeof(float));
  unsigned* hflag;
  cudaHostAlloc(&hflag, 64, cudaHostAllocMapped);
  unsigned* dflag;
  cudaHostGetDevicePointer(&dflag, hflag, 0);
  *hflag = 0;
  cudaStream_t s;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  const int n = 2000;
  for (int spin : {0, 4000}) {   // ~2 us and ~12 us kernels
    std::vector<double> a(n), b(n), c(n);
    unsigned seq = 0;
    for (int i = 0; i < n + 200; i++) {  // (a) stream synchronise
      double t0 = now_us();
      k_work<<<1, 256, 0, s>>>(x, nullptr, 0, spin);
      cudaStreamSynchronize(s);
      if (i >= 200) a[i - 200] = now_us() - t0;
    }
    for (int i = 0; i < n + 200; i++) {  // (b) spin on the mapped flag, no synchronise call at all
      double t0 = now_us();
      k_work<<<1, 256, 0, s>>>(x, dflag, ++seq, spin);
      while (*(volatile unsigned*)hflag != seq) {}
      if (i >= 200) b[i - 200] = now_us() - t0;
    }
    cudaStreamSynchronize(s);
    for (int i = 0; i < n + 200; i++) {  // (c) spin on cudaStreamQuery
      double t0 = now_us();
      k_work<<<1, 256, 0, s>>>(x, nullptr, 0, spin);
      while (cudaStreamQuery(s) == cudaErrorNotReady) {}
      if (i >= 200) c[i - 200] = now_us() - t0;
    }
    std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end()); std::sort(c.begin(), c.end());
    printf("spin %5d: launch + cudaStreamSynchronize p50 %.1f p99 %.1f us | launch + flag spin p50 %.1f p99 %.1f us | launch + cudaStreamQuery spin p50 %.1f p99 %.1f us\n",
           spin, a[n / 2], a[n * 99 / 100], b[n / 2], b[n * 99 / 100], c[n / 2], c[n * 99 / 100]);
  }
  return 0;
}
