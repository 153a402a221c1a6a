// tc_probe.cu -- standalone check of the tcgen05 building blocks used by the MIMO kernel (development tool).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe tools/tc_probe.cu && timeout 60 /tmp/tc_probe
//
// C[128 x 128] = A[128 x K] * B[K x 128] in one CTA: A staged K-major, B staged MN-major, both in the
// no-swizzle canonical core-matrix layout, kind::tf32, accumulator in TMEM, read back with tcgen05.ld.
// Prints the max error against a host reference on tf32-representable inputs (exact products).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

static constexpr int M = 128, N = 128, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                               int* __restrict__ status, int mode) {
  // A: [M/8 m-cores][..] layout: k-core major: offset(r, c) = (c/4) * (M/8)*128 + (r/8)*128 + (r%8)*16 + (c%4)*4  bytes
  // B: offset(k, n) = (k/8) * (N/4)*128 + (n/4)*128 + (k%8)*16 + (n%4)*4 bytes
  __shared__ __align__(128) float sA[M * K];
  __shared__ __align__(128) float sB[K * N];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int idx = tid; idx < M * K; idx += 128) {
    int r = idx / K, c = idx % K;
    int off = (c / 4) * (M / 8) * 32 + (r / 8) * 32 + (r % 8) * 4 + (c % 4);
    sA[off] = A[idx];
  }
  for (int idx = tid; idx < K * N; idx += 128) {
    int k = idx / N, n = idx % N;
    int off = (k / 8) * (N / 4) * 32 + (n / 4) * 32 + (k % 8) * 4 + (n % 4);
    if (mode == 1 || mode == 3) off = (k / 4) * (N / 8) * 32 + (n / 8) * 32 + (n % 8) * 4 + (k % 4);  // K-major like A
    sB[off] = B[idx];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // make the generic-proxy smem writes visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (mode == 3) {
    // A operand in TMEM: row = lane, K along columns (one 32-bit column per tf32 element), written with tcgen05.st
    // at columns 128 .. 128+K-1 (the accumulator uses columns 0 .. 127)
    const int row = warp * 32 + (tid & 31);
    for (int c = 0; c < K; c++) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 128u + (uint32_t)c;
      uint32_t v = __float_as_uint(A[row * K + c]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(M >> 4) << 24);
      const uint32_t b0 = smem_u32(sB);
      for (int j = 0; j < K / 8; j++) {
        const uint64_t db = make_desc(b0 + j * 2 * (N / 8) * 128, (N / 8) * 128, 128);
        const uint32_t acc = j > 0 ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
            :
            : "r"(tmem), "r"(tmem + 128u + (uint32_t)(j * 8)), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
  } else if (mode == 2) {
    // no MMA: write a pattern into TMEM with tcgen05.st and read it back below (checks the ld path / addressing)
    for (int c0 = 0; c0 < N; c0 += 32) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      for (int i = 0; i < 32; i++) {
        uint32_t v = __float_as_uint((float)((warp * 32 + (tid & 31)) * 1000 + c0 + i));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + i), "r"(v) : "memory");
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (tid == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
  } else if (tid == 0) {
    // instruction descriptor: c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, a_major K (0) @15,
    // b_major MN (1) @16, n_dim N>>3 @17, m_dim M>>4 @24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((mode == 1 ? 0u : 1u) << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int j = 0; j < K / 8; j++) {
      // A: K-major, LBO = distance between the two 16-byte K halves = (M/8)*128, SBO = distance between 8-row groups = 128
      const uint64_t da = make_desc(a0 + j * 2 * (M / 8) * 128, (M / 8) * 128, 128);
      // B: MN-major, one 8-deep K group per MMA: SBO = distance between 4-wide N groups = 128, LBO = next K group
      const uint64_t db = (mode == 1) ? make_desc(b0 + j * 2 * (N / 8) * 128, (N / 8) * 128, 128)
                                      : make_desc(b0 + j * (N / 4) * 128, (N / 4) * 128, 128);
      const uint32_t acc = j > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          :
          : "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // bounded wait: never hang the GPU box
  bool done = false;
  for (int spin = 0; spin < 2000000 && !done; spin++) done = mbar_try_wait(smem_u32(&bar), 0);
  if (!done) {
    if (tid == 0) status[0] = -1;
  } else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes 32w .. 32w+31 (= rows of C), 128 columns in 4 chunks of 32
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = warp * 32 + (tid & 31);
      for (int i = 0; i < 32; i++) C[row * N + c0 + i] = __uint_as_float(r[i]);
    }
    if (tid == 0) status[0] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main(int argc, char** argv) {
  std::vector<float> A(M * K), B(K * N), C(M * N, 0.f), R(M * N, 0.f);
  srand(1);
  // small integers / 8: exactly representable in tf32, exact fp32 accumulation
  for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.0f;
  for (auto& v : B) v = (float)((rand() % 17) - 8) / 8.0f;
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      double s = 0;
      for (int k = 0; k < K; k++) s += (double)A[m * K + k] * B[k * N + n];
      R[m * N + n] = (float)s;
    }
  float *dA, *dB, *dC;
  int* dS;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dC, C.size() * 4);
  cudaMalloc(&dS, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  int rc = 1;
  for (int mode = 0; mode < 4; mode++) {
    cudaMemset(dS, 0, 4);
    cudaMemset(dC, 0, C.size() * 4);
    k_probe<<<1, 128>>>(dA, dB, dC, dS, mode);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0, nz = 0;
    for (int i = 0; i < M * N; i++) {
      double want = (mode == 2) ? (double)((i / N) * 1000 + (i % N)) : R[i];
      double d = fabs((double)C[i] - want);
      if (d > maxerr) maxerr = d;
      if (d > 1e-4) bad++;
      if (C[i] != 0.f) nz++;
    }
    printf("mode=%d cuda=%s status=%d maxerr=%g bad=%d nonzero=%d  C[0..3]=%g %g %g %g  C[129]=%g R[0..3]=%g %g %g %g R[129]=%g\n", mode,
           cudaGetErrorString(e), st, maxerr, bad, nz, C[0], C[1], C[2], C[3], C[129], R[0], R[1], R[2], R[3], R[129]);
    if (mode == 1 && e == cudaSuccess && st == 1 && bad == 0) rc = 0;
    if (e != cudaSuccess) break;
  }
  return rc;
}
