// mimo_tc.cuh -- MIMO frequency-domain mixdown as a per-bin complex GEMM on the tcgen05 tensor cores.
//
// Y_o[k] = sum_i sum_p H_{o,i}[p][k] * FDL_i[(head - p)][k].  With T block-steps in one
// call this is, for every bin k, a dense complex GEMM
//       Y[o][t] = sum_j  H[o][j] * X[j][t],     j = (i, p),  X[(i,p)][t] = FDL_i[slot(t - p)]
// (M = n_out, K = n_in * P, N = T).  The complex product is embedded in a real one so that ONE accumulator
// tile fills the 128 TMEM lanes:
//       rows 0..63   (re of output o):  [ Hr, -Hi ] . [ Xr, Xi ]
//       rows 64..127 (im of output o):  [ Hi,  Hr ] . [ Xr, Xi ]
// i.e. A = 128 x 2K (built on the fly from the raw complex spectra), B = 2K x N (the FDL values as stored).
// fp32 accuracy (SNR >= 110 dB) needs more than one TF32 pass: both operands are split v = hi + lo with
// hi = tf32_rna(v), and three MMAs (lo*hi, hi*lo, hi*hi) accumulate into fp32 TMEM tiles (see kTcAccTiles).
//
// Operand layouts in HBM (bin-major, written by the pack kernels below):
//   Hpack[og][k][g][64] float4   g indexes pairs of complex K elements j = 2g, 2g+1;  j = i*P2 + p',
//                                 p' = P2-1-p (partition order reversed so a column of B is a contiguous
//                                 run of the time axis), P2 = power of two >= P, K padded to 16 with zeros
//   Xb[k][i][w] float2           w = t + p': block t - p of this call (negative = history) -> W = P2-1+Tcap
// One CTA owns kTcBins adjacent bins (their 8-byte output writes fill one 32-byte sector in L2) and walks K in chunks
// of 16 complex: all 256 threads convert the chunk (hi/lo split, sign/swap expansion) into the canonical
// no-swizzle K-major core-matrix layout in shared memory and arrive on the stage's "full" mbarrier; a ninth
// warp waits for it and issues the 12 tcgen05.mma of the chunk; tcgen05.commit on the stage's "empty" mbarrier
// releases it for re-use (4 stages).  No block-wide barrier in the main loop.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bbx {

static constexpr int kTcProducers = 256;            // warps 0..7: operand conversion + epilogue
static constexpr int kTcThreads = kTcProducers + 32;  // warp 8: one elected lane issues the MMAs
static constexpr int kTcBins = 4;      // adjacent bins per CTA
static constexpr int kTcStages = 4;
static constexpr int kTcChunk = 16;    // complex K elements per stage = 32 tf32 = 4 MMA k-steps
static constexpr int kTcRows = 128;    // accumulator rows: 64 outputs x (re, im)
static constexpr int kTcNmax = 64;     // columns (block-steps) per accumulator tile
static constexpr uint32_t kTcAHalf = kTcRows * kTcChunk * 2 * 4;          // 16 KB: A_hi (then A_lo)
static constexpr uint32_t kTcBHalf = kTcNmax * kTcChunk * 2 * 4;          // 8 KB:  B_hi (then B_lo)
static constexpr uint32_t kTcStageBytes = 2 * kTcAHalf + 2 * kTcBHalf;    // 48 KB
static constexpr uint32_t kTcTileBytes = kTcNmax * 2 * 64 * 4;            // 32 KB: epilogue tile [t][re/im][o]
static constexpr uint32_t kTcSmemBytes = kTcStages * kTcStageBytes + kTcTileBytes + 128;  // + barriers
// TMEM: two accumulator sets (bins alternate, so a set drains while the other fills) of four 64-column tiles.
// The tensor core truncates the fp32 accumulator on every MMA, a bias that grows with the number of sequential
// accumulations (one tile for everything: 109 dB SNR at K = 512 complex).  The hi*hi products therefore rotate
// over three tiles (k-step mod 3) and the small lo*hi / hi*lo terms have their own tile, so the dominant sums see
// K/12 accumulations instead of 3K/4; the epilogue adds the four tiles in fp32 round-to-nearest.
static constexpr uint32_t kTcAccTiles = 4;
static constexpr uint32_t kTcTmemCols = 512;                              // 2 x 4 x 64
static constexpr uint32_t kTcMaxK = 1024;                                 // complex K verified against the tolerance

struct MimoTcArgs {
  const float4* hpack;
  const float2* xb;
  float2* ypart;     // [t][slot_stride][B], slot = output
  int* status;       // set non-zero when a barrier wait times out (never hang the device)
  uint32_t B, n_in, n_out, P2log, G /* float4 K groups = Kc/2 */, W, T, N /* 16, 32 or 64 */, Nlog, slot_stride;
};

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major: LBO = byte distance between the two 16-byte K halves
// of one MMA k-step, SBO = byte distance between 8-row groups (tools/tc_probe.cu validates this encoding)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
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

// bounded wait: returns false after ~1e6 polls
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 20); spin++)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// v = hi + lo with hi = v rounded to TF32 (nearest, ties away: add half an ulp of the 10-bit mantissa to the
// sign-magnitude pattern and clear the 13 low bits; cvt.rna.tf32.f32 compiles to a ~6-instruction sequence with
// NaN handling on sm_100a, this is two).  v - hi is exact; the remainder gets the half-ulp only: the tensor core
// ignores the 13 low bits of a TF32 operand, which completes the rounding.
__device__ __forceinline__ void split(float v, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
  lo = __uint_as_float(__float_as_uint(v - hi) + 0x1000u);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ void st_shared4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_shared2(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

}  // namespace tc

__global__ void __launch_bounds__(kTcThreads, 1) k_mimo_tc(MimoTcArgs a) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  using namespace tc;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t kb = blockIdx.x * kTcBins, og = blockIdx.y, t0 = blockIdx.z * a.N;
  const uint32_t N = a.N, Nlog = a.Nlog;
  const uint32_t smem0 = smem_u32(tc_smem);
  float* tile = reinterpret_cast<float*>(tc_smem + kTcStages * kTcStageBytes);
  // mbarriers: full[NST] (256 producer arrivals), empty[NST] (tcgen05.commit), acc_full[2] (commit), acc_empty[2] (256)
  const uint32_t bar_full = smem0 + kTcStages * kTcStageBytes + kTcTileBytes;
  const uint32_t bar_empty = bar_full + 8 * kTcStages;
  const uint32_t bar_accf = bar_empty + 8 * kTcStages;
  const uint32_t bar_acce = bar_accf + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_smem + kTcStages * kTcStageBytes + kTcTileBytes + 8 * (2 * kTcStages + 4));

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTcStages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8 * s), "n"(kTcProducers));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_empty + 8 * s));
    }
#pragma unroll
    for (int b = 0; b < 2; b++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_accf + 8 * b));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_acce + 8 * b), "n"(kTcProducers));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTcTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t nchunk = a.G / (kTcChunk / 2);
  bool ok = true;

  if (warp == kTcProducers / 32) {
    // ================= MMA issuer: one lane =================
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 @17, M >> 4 @24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
      uint32_t s = 0, ph = 0;
      for (uint32_t j = 0; j < (uint32_t)kTcBins; j++) {
        // accumulator set j & 1: tiles 0..2 take the hi*hi products by k-step mod 3, tile 3 the small terms
        const uint32_t d = tmem + (j & 1) * kTcAccTiles * kTcNmax;
        if (j >= 2 && ok && !mbar_wait(bar_acce + 8 * (j & 1), ((j >> 1) - 1) & 1)) ok = false;  // set drained
        uint32_t rot = 0, ks = 0;
        for (uint32_t c = 0; c < nchunk; c++) {
          if (ok && !mbar_wait(bar_full + 8 * s, ph)) ok = false;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sA = smem0 + s * kTcStageBytes, sB = sA + 2 * kTcAHalf;
#pragma unroll
          for (int k8 = 0; k8 < kTcChunk / 4; k8++) {
            const uint64_t a_hi = make_desc(sA + k8 * 2 * (kTcRows * 16), kTcRows * 16, 128);
            const uint64_t a_lo = make_desc(sA + kTcAHalf + k8 * 2 * (kTcRows * 16), kTcRows * 16, 128);
            const uint64_t b_hi = make_desc(sB + k8 * 2 * (N * 16), N * 16, 128);
            const uint64_t b_lo = make_desc(sB + kTcBHalf + k8 * 2 * (N * 16), N * 16, 128);
            mma_tf32(d + 3 * kTcNmax, a_lo, b_hi, idesc, ks ? 1u : 0u);
            mma_tf32(d + 3 * kTcNmax, a_hi, b_lo, idesc, 1u);
            mma_tf32(d + rot * kTcNmax, a_hi, b_hi, idesc, ks >= 3u ? 1u : 0u);
            rot = rot == 2 ? 0 : rot + 1;
            ks++;
          }
          commit(bar_empty + 8 * s);
          if (c + 1 == nchunk) commit(bar_accf + 8 * (j & 1));
          if (++s == kTcStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      if (!ok) atomicExch(a.status, 2);
    }
  } else {
    // ================= producers (256 threads) =================
    // A: this thread owns output ao of the group and K groups agq, agq + 4 of every chunk; the packed operand is
    // linear in (bin, chunk): 512 float4 per chunk
    const uint32_t ao = tid & 63, agq = tid >> 6;
    const float4* hp = a.hpack + ((uint64_t)(og * a.B + kb) * a.G) * 64 + (uint64_t)agq * 64 + ao;
    const uint32_t offA = agq * (kTcRows * 16) + ao * 16;
    // B: items e = tid + 256 r -> (pair member, column t, K group)
    const uint32_t nbr = N >> 4;
    const uint32_t P2m = (1u << a.P2log) - 1;
    uint32_t offB[4], jlB[4], tB[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t e = tid + 256 * r;
      const uint32_t kg = e >> (1 + Nlog);
      tB[r] = (e >> 1) & (N - 1);
      jlB[r] = (kg << 1) | (e & 1);
      offB[r] = kg * (N * 16) + tB[r] * 16 + (e & 1) * 8;
    }
    const float2* xk0 = a.xb + (uint64_t)kb * a.n_in * a.W + t0;
    const uint64_t xbin = (uint64_t)a.n_in * a.W;

    float4 ra[2][2];
    float2 rb[2][4];
    uint32_t lj = 0, lc = 0;  // (bin, chunk) of the next load
    auto load_chunk = [&](float4(&qa)[2], float2(&qb)[4]) {
      if (lj < (uint32_t)kTcBins) {
        qa[0] = ld_stream4(hp);
        qa[1] = ld_stream4(hp + 4 * 64);
        hp += 8 * 64;
        const float2* xk = xk0 + lj * xbin;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          qb[r] = make_float2(0.f, 0.f);
          if ((uint32_t)r < nbr) {
            const uint32_t jj = lc * kTcChunk + jlB[r], i = jj >> a.P2log, pp = jj & P2m;
            if (i < a.n_in) qb[r] = __ldg(xk + (uint64_t)i * a.W + tB[r] + pp);
          }
        }
        if (++lc == nchunk) {
          lc = 0;
          lj++;
        }
      }
    };

    // drain the accumulator set of bin j: sum its four tiles, pair (re, im) through the smem tile, store 8 bytes
    // per (t, output).  The four bins of a CTA fill one 32-byte sector within microseconds: L2 merges the writes.
    auto drain_bin = [&](uint32_t j) {
      const uint32_t b = j & 1;
      asm volatile("bar.sync 1, %0;" ::"n"(kTcProducers) : "memory");  // the previous drain's readers are done with the tile
      if (ok && !mbar_wait(bar_accf + 8 * b, (j >> 1) & 1)) ok = false;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t q = warp & 3, h = warp >> 2;
      const uint32_t m = 32 * q + lane, cc = m >> 6, o = m & 63;
      for (uint32_t cg = h; cg < (N >> 4); cg += 2) {
        uint32_t r[kTcAccTiles][16];
#pragma unroll
        for (uint32_t z = 0; z < kTcAccTiles; z++) {
          const uint32_t taddr = tmem + ((32 * q) << 16) + (b * kTcAccTiles + z) * kTcNmax + cg * 16;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[z][0]), "=r"(r[z][1]), "=r"(r[z][2]), "=r"(r[z][3]), "=r"(r[z][4]), "=r"(r[z][5]), "=r"(r[z][6]),
                "=r"(r[z][7]), "=r"(r[z][8]), "=r"(r[z][9]), "=r"(r[z][10]), "=r"(r[z][11]), "=r"(r[z][12]), "=r"(r[z][13]),
                "=r"(r[z][14]), "=r"(r[z][15])
              : "r"(taddr));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 16; u++) {
          const uint32_t t = cg * 16 + u;
          const float v = ((__uint_as_float(r[0][u]) + __uint_as_float(r[1][u])) + __uint_as_float(r[2][u])) + __uint_as_float(r[3][u]);
          tile[(t * 2 + cc) * 64 + o] = v;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_acce + 8 * b);  // this thread's reads of the set are complete
      asm volatile("bar.sync 1, %0;" ::"n"(kTcProducers) : "memory");
      for (uint32_t idx = tid; idx < N * 64; idx += kTcProducers) {
        const uint32_t oo = idx & 63, t = idx >> 6;
        const uint32_t og_o = og * 64 + oo, tt = t0 + t;
        if (tt < a.T && og_o < a.n_out)
          a.ypart[((uint64_t)tt * a.slot_stride + og_o) * a.B + kb + j] = make_float2(tile[(t * 2) * 64 + oo], tile[(t * 2 + 1) * 64 + oo]);
      }
    };

    uint32_t s = 0, ph = 0, it = 0;
    auto produce = [&](float4(&qa)[2], float2(&qb)[4], uint32_t j) {
      const uint32_t sA = smem0 + s * kTcStageBytes, sB = sA + 2 * kTcAHalf;
      if (it >= (uint32_t)kTcStages) {
        // the MMAs that read this stage kTcStages chunks ago have completed
        if (ok && !mbar_wait(bar_empty + 8 * s, ph ^ 1)) ok = false;
      }
      const bool bin0 = (kb + j) == 0;  // packed bin 0 = (DC, Nyquist): two real products, no cross terms
      // ---- A: raw (a0, b0, a1, b1) = two complex of output ao -> rows ao (re) and 64 + ao (im), hi and lo ----
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const float4 v = qa[q];
        float ah0, al0, bh0, bl0, ah1, al1, bh1, bl1;
        split(v.x, ah0, al0);
        split(v.y, bh0, bl0);
        split(v.z, ah1, al1);
        split(v.w, bh1, bl1);
        const uint32_t off = offA + q * 4 * (kTcRows * 16);
        if (!bin0) {
          st_shared4(sA + off, ah0, -bh0, ah1, -bh1);
          st_shared4(sA + off + 64 * 16, bh0, ah0, bh1, ah1);
          st_shared4(sA + kTcAHalf + off, al0, -bl0, al1, -bl1);
          st_shared4(sA + kTcAHalf + off + 64 * 16, bl0, al0, bl1, al1);
        } else {
          st_shared4(sA + off, ah0, 0.f, ah1, 0.f);
          st_shared4(sA + off + 64 * 16, 0.f, bh0, 0.f, bh1);
          st_shared4(sA + kTcAHalf + off, al0, 0.f, al1, 0.f);
          st_shared4(sA + kTcAHalf + off + 64 * 16, 0.f, bl0, 0.f, bl1);
        }
      }
      // ---- B: one complex of column t -> 8 bytes of the K-major tile, hi and lo ----
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if ((uint32_t)r < nbr) {
          float xh, xl, yh, yl;
          split(qb[r].x, xh, xl);
          split(qb[r].y, yh, yl);
          st_shared2(sB + offB[r], xh, yh);
          st_shared2(sB + kTcBHalf + offB[r], xl, yl);
        }
      }
      load_chunk(qa, qb);  // refill this register set: the chunk two iterations ahead
      // generic-proxy writes -> visible to the tensor core (async proxy), then hand the stage to the issuer
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_full + 8 * s);
      if (++s == kTcStages) {
        s = 0;
        ph ^= 1;
      }
      it++;
    };

    load_chunk(ra[0], rb[0]);
    load_chunk(ra[1], rb[1]);
    const uint32_t cdrain = nchunk > 1 ? 1u : 0u;  // chunk of bin j after which bin j-1 is drained
    uint32_t par = 0;
    for (uint32_t j = 0; j < (uint32_t)kTcBins; j++)
      for (uint32_t c = 0; c < nchunk; c++) {
        if (par == 0) produce(ra[0], rb[0], j);
        else produce(ra[1], rb[1], j);
        par ^= 1;
        if (j > 0 && c == cdrain) drain_bin(j - 1);
      }
    drain_bin(kTcBins - 1);
    if (!ok && tid == 0) atomicExch(a.status, 1);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTcTmemCols));
}

// Xb[k][i][w] = FDL[i][slot of block (w - (P2-1))][k] for w < P2-1+T, zero for the padding columns.
// 32 x 32 tile transpose (k <-> w) through shared memory, one input per blockIdx.z.
__global__ void __launch_bounds__(256) k_mimo_pack_x(const float2* __restrict__ fdl, float2* __restrict__ xb, uint32_t B,
                                                     uint32_t R, uint32_t head, uint32_t n_in, uint32_t P2, uint32_t T,
                                                     uint32_t W) {
  __shared__ float2 tile[32][33];
  const uint32_t k0 = blockIdx.x * 32, w0 = blockIdx.y * 32, i = blockIdx.z;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t wl = ty + 8 * r, w = w0 + wl;
    float2 v = make_float2(0.f, 0.f);
    if (w < P2 - 1 + T) {
      long long blk = (long long)head + (long long)w - (long long)(P2 - 1);
      long long sl = blk % (long long)R;
      if (sl < 0) sl += R;
      v = fdl[((uint64_t)i * R + (uint64_t)sl) * B + k0 + tx];
    }
    tile[wl][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t kl = ty + 8 * r, w = w0 + tx;
    if (w < W) xb[((uint64_t)(k0 + kl) * n_in + i) * W + w] = tile[tx][kl];
  }
}

// Hpack[og][k][g][o] = (H_{o,i0}[p0][k], H_{o,i1}[p1][k]) for the complex K indices j = 2g, 2g+1 (j = i*P2 + p',
// p = P2-1-p'); zero where the filter is null / shorter / beyond n_in.  Runs when the filter matrix changes.
__global__ void __launch_bounds__(256) k_mimo_pack_h(const float2* const* __restrict__ ftab, const uint32_t* __restrict__ fparts,
                                                     float4* __restrict__ hpack, uint32_t B, uint32_t n_in, uint32_t n_out,
                                                     uint32_t P2log, uint32_t G) {
  __shared__ float2 tile[32][129];  // [k][o*2 + jj]
  const uint32_t k0 = blockIdx.x * 32, g = blockIdx.y, og = blockIdx.z;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t P2 = 1u << P2log;
  for (uint32_t pr = warp; pr < 128; pr += 8) {  // (o, jj) pairs, lanes over k
    const uint32_t o = pr >> 1, jj = pr & 1;
    const uint32_t j = 2 * g + jj, i = j >> P2log, p = P2 - 1 - (j & (P2 - 1));
    const uint32_t og_o = og * 64 + o;
    float2 v = make_float2(0.f, 0.f);
    if (og_o < n_out && i < n_in) {
      const float2* H = ftab[(uint64_t)og_o * n_in + i];
      if (H && p < fparts[(uint64_t)og_o * n_in + i]) v = H[(uint64_t)p * B + k0 + lane];
    }
    tile[lane][pr] = v;
  }
  __syncthreads();
  for (uint32_t kl = warp; kl < 32; kl += 8)
    for (uint32_t o = lane; o < 64; o += 32) {
      const float2 u = tile[kl][2 * o], w = tile[kl][2 * o + 1];
      hpack[(((uint64_t)og * B + k0 + kl) * G + g) * 64 + o] = make_float4(u.x, u.y, w.x, w.y);
    }
}

}  // namespace bbx
