// ffma2_probe.cu -- what FP32 rate can a stream of packed FFMA2 (fma.rn.f32x2) reach on sm_100a with the operand pattern of
// k_fdl_mac_tb (16 float2 accumulators per thread, one broadcast multiplier pair per step)?  Development tool.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_probe tools/ffma2_probe.cu && build/ffma2_probe
#include <cuda_runtime.h>
#include <stdio.h>

template <int NACC, bool PACKED>
__global__ void __launch_bounds__(256) k_probe(float2* out, int iters, float2 h0, float2 x0) {
  float2 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 h = h0, x = x0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (PACKED) {
        acc[i] = __ffma2_rn(h, make_float2(x.x, x.x), acc[i]);
        acc[i] = __ffma2_rn(make_float2(-h.y, h.x), make_float2(x.y, x.y), acc[i]);
      } else {
        acc[i].x = fmaf(h.x, x.x, acc[i].x);
        acc[i].x = fmaf(-h.y, x.y, acc[i].x);
        acc[i].y = fmaf(h.y, x.x, acc[i].y);
        acc[i].y = fmaf(h.x, x.y, acc[i].y);
      }
    }
    h.x += 1e-7f;  // keep the loop from being hoisted
    x.y -= 1e-7f;
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NACC; i++) {
    s.x += acc[i].x;
    s.y += acc[i].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, bool PACKED>
void run(const char* name, int ctas_per_sm) {
  const int iters = 4096, grid = 148 * ctas_per_sm;
  float2* out;
  cudaMalloc(&out, sizeof(float2) * grid * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_probe<NACC, PACKED><<<grid, 256>>>(out, 16, make_float2(1.0001f, 0.0001f), make_float2(0.9999f, 0.0002f));
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_probe<NACC, PACKED><<<grid, 256>>>(out, iters, make_float2(1.0001f, 0.0001f), make_float2(0.9999f, 0.0002f));
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 8.0 * NACC * (double)iters * grid * 256;
  printf("%-28s ctas/SM %d  %8.3f ms  %7.2f TFLOP/s\n", name, ctas_per_sm, ms, flops / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  for (int c = 1; c <= 4; c *= 2) {
    run<16, true>("FFMA2, 16 accumulators", c);
    run<16, false>("FFMA,  16 accumulators", c);
    run<8, true>("FFMA2, 8 accumulators", c);
  }
  return 0;
}
