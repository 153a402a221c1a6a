// kernels_mac.cuh -- the FDL multiply-accumulate kernels: streaming (HBM-bound), time-batched (FP32-bound), the Nyquist side
// sums, and the FP32 roofline probe that measures the ceiling of the time-batched kernel's instruction mix.
#pragma once

#include "kernels_common.cuh"
#include "mac_common.cuh"

namespace bbx {

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  // read-once data: bypass L1 allocation
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_policy(const float4* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float2 ld_stream2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}

// ---- streaming form: one launch covers nt block-steps, every step re-streams H and the FDL ----
template <int U, int THREADS, int OCC, bool POLICY>
__global__ void __launch_bounds__(THREADS, OCC)
k_fdl_mac(const MacSeg* __restrict__ segs, const uint32_t* __restrict__ cta_seg_begin, const float4* __restrict__ fdl,
          float4* __restrict__ ypart, float* __restrict__ nyq_part, uint32_t halfB, uint32_t R, uint32_t head0,
          uint32_t t0, uint32_t slot_stride, float l2_keep, int policy_x) {
  const uint32_t t = t0 + blockIdx.z;
  const uint32_t head = (head0 + t) % R;
  const uint32_t col = blockIdx.y * THREADS + threadIdx.x;
  const uint32_t sb = cta_seg_begin[blockIdx.x], se = cta_seg_begin[blockIdx.x + 1];
  uint64_t pol = 0;
  if (POLICY) {
    // keep a fixed fraction of the lines resident in L2 across block-steps (the same H / FDL addresses are
    // re-read every step), stream the rest with evict-first so they do not displace the resident set
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(l2_keep));
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float nacc = 0.f;  // Nyquist sum of the row's first slot; only column 0's copy is meaningful
  for (uint32_t si = sb; si < se; si++) {
    const MacSeg sg = segs[si];
    if (sg.flags & 1u) {
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      nacc = 0.f;
    }
    const float4* hp = sg.H + (uint64_t)sg.p0 * halfB + col;
    const float4* xbase = fdl + (uint64_t)sg.fdl_ch * R * halfB + col;
    int slot = (int)head - (int)sg.p0;  // p0 < R
    if (slot < 0) slot += (int)R;
    // U rows per iteration, all 2U loads issued before the first FMA; rows past the end of a short segment are
    // predicated off and contribute h = x = 0 (acc += 0 exactly), so short filters keep their loads in flight
    for (uint32_t p = 0; p < sg.np; p += U) {
      float4 h[U], x[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        int s = slot - u;
        s += (s >> 31) & (int)R;  // ring wrap, once per row
        h[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p + u < sg.np) {
          if (POLICY) {
            h[u] = ld_policy(hp + (uint64_t)u * halfB, pol);
            x[u] = policy_x ? ld_policy(xbase + (uint64_t)s * halfB, pol) : __ldg(xbase + (uint64_t)s * halfB);
          } else {
            h[u] = ld_stream(hp + (uint64_t)u * halfB);
            x[u] = __ldg(xbase + (uint64_t)s * halfB);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        cmac(acc.x, acc.y, h[u].x, h[u].y, x[u].x, x[u].y);
        cmac(acc.z, acc.w, h[u].z, h[u].w, x[u].z, x[u].w);
        nacc = fmaf(h[u].y, x[u].y, nacc);
      }
      hp += (uint64_t)U * halfB;
      slot -= U;
      if (slot < 0) slot += (int)R;
    }
    if (sg.flags & 2u) {
      ypart[((uint64_t)blockIdx.z * slot_stride + sg.slot) * halfB + col] = acc;
      if (col == 0) nyq_part[(uint64_t)blockIdx.z * slot_stride + sg.slot] = nacc;
    }
  }
}

// ---- time-batched form: one CTA produces TT consecutive block-steps of its row range at once ----
// Y_t[k] = sum_p H[p][k] X[s_t - p][k] for t = t_base .. t_base+TT-1 is a length-P FIR along the block axis:
// H[p] is loaded once and applied to TT outputs, and the TT FDL rows it meets slide by one row per
// partition, so the rows live in a register window rotated by static indexing (the p loop is unrolled TT
// times).  Per partition a thread loads one complex of H and one of the FDL and issues 4*TT FMAs
// (8 FMA per byte at TT = 32): the MAC becomes FP32-bound instead of HBM-bound.  Same plan, same per-output
// FMA order as k_fdl_mac, hence bit-identical results.
__device__ __forceinline__ void cp_async8(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Operands are staged through shared memory with cp.async: every thread copies the H and FDL values of its own
// column NST-1 partitions ahead and reads them back itself (no block barrier), and cp.async.wait_group gives
// the "at most N groups pending" wait that register-target loads cannot express (their scoreboards only count to
// zero, which collapses a deep software pipeline to a depth of one; see profiles/r01 notes in DESIGN.md).
template <int TT, int THREADS, int NST>
__global__ void __launch_bounds__(THREADS, (TT <= 16) ? 2 : 1)
k_fdl_mac_tb(const MacSeg* __restrict__ segs, const uint32_t* __restrict__ cta_seg_begin, const float2* __restrict__ fdl,
             float2* __restrict__ ypart, uint32_t B, uint32_t R, uint32_t head0, uint32_t t0, uint32_t nt,
             uint32_t ncoltiles, uint32_t slot_stride) {
  static_assert(TT % NST == 0, "stage count must divide the tile so that stage indices are static");
  __shared__ float2 stage[NST][2][THREADS];
  // blockIdx.x enumerates (column tile, t tile) so the CTAs that share H / FDL rows run in the same wave
  const uint32_t coltile = blockIdx.x % ncoltiles, ttile = blockIdx.x / ncoltiles;
  const uint32_t tbase = ttile * TT;                  // first block-step of this tile, relative to t0
  const uint32_t s0 = (head0 + t0 + tbase) % R;       // its FDL slot
  const uint32_t col = coltile * THREADS + threadIdx.x;
  const uint32_t sb = cta_seg_begin[blockIdx.y], se = cta_seg_begin[blockIdx.y + 1];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&stage[0][0][threadIdx.x]);
  constexpr uint32_t kStageBytes = 2 * THREADS * sizeof(float2);
  constexpr uint32_t kXOff = THREADS * sizeof(float2);
  const uint32_t row_bytes = B * (uint32_t)sizeof(float2);
  float2 acc[TT];
#pragma unroll
  for (int i = 0; i < TT; i++) acc[i] = make_float2(0.f, 0.f);
  for (uint32_t si = sb; si < se; si++) {
    const MacSeg sg = segs[si];
    if (sg.flags & 1u) {
#pragma unroll
      for (int i = 0; i < TT; i++) acc[i] = make_float2(0.f, 0.f);
    }
    // byte pointers: one 32x32+64 multiply-add per address
    const char* hp = reinterpret_cast<const char*>(reinterpret_cast<const float2*>(sg.H) + (uint64_t)sg.p0 * B + col);
    const char* xb = reinterpret_cast<const char*>(fdl + (uint64_t)sg.fdl_ch * R * B + col);
    // FDL row met by output i at segment step q: base + i - q (mod R), base = s0 - p0
    uint32_t base = s0 + R - (sg.p0 % R);
    if (base >= R) base -= R;
    uint32_t prow = base;  // FDL row of the next step to be staged
    // stage the first NST-1 steps (one commit group per step, empty past the end so the count stays uniform)
#pragma unroll
    for (int j = 0; j < NST - 1; j++) {
      if ((uint32_t)j < sg.np) {
        cp_async8(sbase + j * kStageBytes, hp + (uint64_t)j * row_bytes);
        cp_async8(sbase + j * kStageBytes + kXOff, xb + (uint64_t)prow * row_bytes);
        prow = prow ? prow - 1 : R - 1;
      }
      cp_async_commit();
    }
    float2 W[TT];  // W[e mod TT] = row base + e, e = i - q
#pragma unroll
    for (int e = 1; e < TT; e++) {
      uint32_t r = base + e;
      if (r >= R) r -= R;
      W[e] = __ldg(reinterpret_cast<const float2*>(xb + (uint64_t)r * row_bytes));
    }
    W[0] = make_float2(0.f, 0.f);
    uint32_t qb = 0;
    // fast path: whole groups of TT steps whose look-ahead stays inside the segment, no per-step checks
    for (; qb + TT + NST - 1 <= sg.np; qb += TT) {
#pragma unroll
      for (int u = 0; u < TT; u++) {
        const uint32_t qn = qb + u + NST - 1;
        const int sn = (u + NST - 1) % NST;
        cp_async8(sbase + sn * kStageBytes, hp + (uint64_t)qn * row_bytes);
        cp_async8(sbase + sn * kStageBytes + kXOff, xb + (uint64_t)prow * row_bytes);
        prow = prow ? prow - 1 : R - 1;
        cp_async_commit();
        cp_async_wait<NST - 1>();  // step qb + u has landed
        const float2 h = stage[u % NST][0][threadIdx.x];
        W[(TT - u) % TT] = stage[u % NST][1][threadIdx.x];
#pragma unroll
        for (int i = 0; i < TT; i++) cmac_x2(acc[i], h, W[(i - u + TT) % TT]);
      }
    }
    // tail: same steps with bound checks
    for (; qb < sg.np; qb += TT) {
#pragma unroll
      for (int u = 0; u < TT; u++) {
        const uint32_t q = qb + u;
        if (q < sg.np) {
          const uint32_t qn = q + NST - 1;
          const int sn = (u + NST - 1) % NST;
          if (qn < sg.np) {
            cp_async8(sbase + sn * kStageBytes, hp + (uint64_t)qn * row_bytes);
            cp_async8(sbase + sn * kStageBytes + kXOff, xb + (uint64_t)prow * row_bytes);
            prow = prow ? prow - 1 : R - 1;
          }
          cp_async_commit();
          cp_async_wait<NST - 1>();
          const float2 h = stage[u % NST][0][threadIdx.x];
          W[(TT - u) % TT] = stage[u % NST][1][threadIdx.x];
#pragma unroll
          for (int i = 0; i < TT; i++) cmac_x2(acc[i], h, W[(i - u + TT) % TT]);
        }
      }
    }
    cp_async_wait<0>();
    if (sg.flags & 2u) {
#pragma unroll
      for (int i = 0; i < TT; i++)
        if (tbase + i < nt) ypart[((uint64_t)(tbase + i) * slot_stride + sg.slot) * B + col] = acc[i];
    }
  }
}

// Nyquist sums next to the time-batched MAC: N[t][run] = sum over the run's rows of Nqh[p] * Nqx[s_t - p] (the
// imaginary parts of column 0), p ascending, one fma per row -- the same sequence the streaming kernel runs inline.
// One warp per (row range, tile of 32 block-steps), lane = block-step.
__global__ void __launch_bounds__(32)
k_nyq_mac(const MacSeg* __restrict__ segs, const uint32_t* __restrict__ cta_seg_begin, const float2* __restrict__ fdl,
          float* __restrict__ nyq_part, uint32_t B, uint32_t R, uint32_t head0, uint32_t t0, uint32_t nt, uint32_t slot_stride) {
  const uint32_t tq = blockIdx.y * 32 + threadIdx.x;
  const bool active = tq < nt;
  const uint32_t t = active ? tq : nt - 1;
  const uint32_t s_t = (head0 + t0 + t) % R;
  const uint32_t sb = cta_seg_begin[blockIdx.x], se = cta_seg_begin[blockIdx.x + 1];
  float acc = 0.f;
  for (uint32_t si = sb; si < se; si++) {
    const MacSeg sg = segs[si];
    if (sg.flags & 1u) acc = 0.f;
    const float2* hcol = reinterpret_cast<const float2*>(sg.H) + (uint64_t)sg.p0 * B;
    const float2* xb = fdl + (uint64_t)sg.fdl_ch * R * B;
    uint32_t row = s_t + R - (sg.p0 % R);
    if (row >= R) row -= R;
#pragma unroll 16
    for (uint32_t p = 0; p < sg.np; p++) {
      const float h = __ldg(&hcol[(uint64_t)p * B]).y;
      const float x = __ldg(&xb[(uint64_t)row * B]).y;
      acc = fmaf(h, x, acc);
      row = row ? row - 1 : R - 1;
    }
    if ((sg.flags & 2u) && active) nyq_part[(uint64_t)t * slot_stride + sg.slot] = acc;
  }
}

// FP32 roofline probe (measurement hook): nothing but packed FMAs on 16 float2 accumulators per thread, the operand
// pattern of k_fdl_mac_tb's inner loop.  Its rate is the FP32 ceiling this GPU reaches under its power / clock limits
// (the nominal 148 x 128 x 2 x 1965 MHz is not reachable: bench.py reports both).
__global__ void __launch_bounds__(256) k_fp32_probe(float2* out, int iters, float2 h0, float2 x0) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 h = h0, x = x0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) cmac_x2(acc[i], h, x);
    h.x += 1e-7f;
    x.y -= 1e-7f;
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    s.x += acc[i].x;
    s.y += acc[i].y;
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace bbx
