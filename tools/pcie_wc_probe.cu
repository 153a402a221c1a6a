// pcie_wc_probe.cu -- host <-> device copy rates of one C3 step's buffers (16.8 MB each way) from ordinary pinned memory and
// from write-combined pinned memory (cudaHostAllocWriteCombined), one direction alone and both at once on two streams.
// nvcc -O2 -o pcie_wc_probe pcie_wc_probe.cu && ./pcie_wc_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const size_t n = 16777216;
  const int reps = 200;
  void *d_in, *d_out, *h_out;
  CK(cudaMalloc(&d_in, n));
  CK(cudaMalloc(&d_out, n));
  CK(cudaHostAlloc(&h_out, n, cudaHostAllocDefault));
  cudaStream_t s0, s1;
  CK(cudaStreamCreate(&s0));
  CK(cudaStreamCreate(&s1));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int wc = 0; wc < 2; wc++) {
    void* h_in;
    CK(cudaHostAlloc(&h_in, n, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    memset(h_in, 1, n);
    for (int mode = 0; mode < 3; mode++) {  // 0: H2D alone, 1: D2H alone, 2: both
      for (int w = 0; w < 5; w++) {
        if (mode != 1) CK(cudaMemcpyAsync(d_in, h_in, n, cudaMemcpyHostToDevice, s0));
        if (mode != 0) CK(cudaMemcpyAsync(h_out, d_out, n, cudaMemcpyDeviceToHost, s1));
      }
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a, s0));
      CK(cudaStreamWaitEvent(s1, a, 0));
      for (int r = 0; r < reps; r++) {
        if (mode != 1) CK(cudaMemcpyAsync(d_in, h_in, n, cudaMemcpyHostToDevice, s0));
        if (mode != 0) CK(cudaMemcpyAsync(h_out, d_out, n, cudaMemcpyDeviceToHost, s1));
      }
      CK(cudaEventRecord(b, s1));
      CK(cudaStreamWaitEvent(s0, b, 0));
      CK(cudaEventRecord(b, s0));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      printf("%s input, %s: %.4f ms per step, %.1f GB/s each way\n", wc ? "write-combined" : "pinned        ",
             mode == 0 ? "H2D alone" : mode == 1 ? "D2H alone" : "both     ", ms / reps, n / (ms / reps) * 1e-6);
    }
    CK(cudaFreeHost(h_in));
  }
  return 0;
}
