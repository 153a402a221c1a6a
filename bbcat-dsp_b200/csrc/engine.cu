// engine.cu -- the partitioned-convolution engine of libbbx (BlockConvolver / Convolver path).
//
// The reference's BlockConvolver.{h,cpp}, Convolver.{h,cpp}, FFT*.cpp and simd_utils (README:38-51,
// 68-69) are absent from the mounted tree; behaviour follows SURVEY.md 8.A.  Data flow of one
// bbx_process call over T blocks of B frames (all kernels on the engine stream):
//
//   k_pcm_in   interleaved PCM (any SampleFormat_t, LE/BE) -> planar fp32 xin[input][(T+1)B]
//              (block 0 of the row is the previous call's last block = the overlap-save history)
//   k_rfft     window [prev | cur] (2B floats, contiguous in xin) -> packed spectrum -> FDL ring slot
//              FDL[input][(head+t) mod R][B] complex, R >= Pmax + Tmax - 1
//   k_fdl_mac  Y[job] = sum over (term, p) H[term][p] * FDL[input(term)][(head+t-p) mod R]
//              the hot kernel: streams H and FDL rows (8B bytes each, 4 KB at B=512) with 128-bit
//              loads, flattened (job, term, p) row space split evenly over a grid sized to the
//              SM count, per-CTA partial sums written to Ypart (deterministic, no atomics)
//   k_irfft    sum the partials in fixed order -> C2R -> keep samples B..2B-1 -> filter crossfade
//              -> per-stream delay ring ybuf[stream][Rd]
//   k_pcm_out  delayed ring reads (integer or 14-tap fractional, delay crossfade) -> MixSamples-order
//              mixdown -> float -> output SampleFormat_t, interleaved
//
// HBM layout: H per filter [P][B] float2 (1/N folded in), FDL [n_in][R][B] float2, both rows
// 16-byte aligned so one thread owns one float4 column (2 bins) across all partitions.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "fft.cuh"
#include "formats.cuh"
#include "fracsample.cuh"
#include "mimo_tc.cuh"

namespace bbx {

static constexpr uint32_t kNoJob = 0xFFFFFFFFu;
static constexpr uint32_t kSameJob = 0xFFFFFFFEu;  // crossfade a stream with itself (delay-only switch)
static constexpr int kNumSMs = 148;
__host__ __device__ __forceinline__ uint32_t ceil_div_dev(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------
// k_pcm_in
// ------------------------------------------------------------------------------------------
struct PcmInArgs {
  const uint8_t* pcm;
  int fmt;
  int be;
  uint32_t in_channels, n_inputs;
  uint32_t B, T;
  float* xin_cur;         // [n_inputs][xstride]
  const float* xin_prev;  // previous call's buffer
  uint32_t xstride;
  uint32_t prev_off;      // offset of the previous call's last block inside xin_prev rows
  int fast;               // little-endian, base and frame stride aligned to the sample size: typed loads
};

template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_in(PcmInArgs a) {
  __shared__ float tile[32][33];
  const uint32_t f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool is_prev = f0 < a.B;  // B is a multiple of 32: a tile never straddles the boundary
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  if (!is_prev) {
    // phase 1: lanes over channels (contiguous bytes within a frame)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t fl = warp + 8 * i, c = c0 + lane;
      uint32_t frame = f0 + fl - a.B;
      float v = 0.f;
      if (c < a.n_inputs) v = load_as_f32_t<FMT, ACC>(a.pcm + ((uint64_t)frame * a.in_channels + c) * bps);
      tile[fl][lane] = v;
    }
    __syncthreads();
  }
  // phase 2: lanes over frames (contiguous floats of one planar row)
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t cl = warp + 8 * i, c = c0 + cl;
    if (c >= a.n_inputs) continue;
    uint32_t f = f0 + lane;
    float v = is_prev ? a.xin_prev[(uint64_t)c * a.xstride + a.prev_off + f] : tile[lane][cl];
    a.xin_cur[(uint64_t)c * a.xstride + f] = v;
  }
}

// Same transpose with 128-frame tiles (B % 128 == 0): 16 independent loads per thread are in flight before the first
// shared-memory store (the 32-frame kernel above is bound by the latency of its 4), a quarter of the CTAs.
template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_in128(PcmInArgs a) {
  __shared__ float tile[128][33];
  const uint32_t f0 = blockIdx.x * 128, c0 = blockIdx.y * 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool is_prev = f0 < a.B;  // B is a multiple of 128: a tile never straddles the boundary
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  if (!is_prev) {
    // phase 1: lanes over channels (contiguous bytes within a frame)
    float v[16];
    const uint32_t c = c0 + lane;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t frame = f0 + warp + 8 * i - a.B;
      v[i] = (c < a.n_inputs) ? load_as_f32_t<FMT, ACC>(a.pcm + ((uint64_t)frame * a.in_channels + c) * bps) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) tile[warp + 8 * i][lane] = v[i];
    __syncthreads();
  }
  // phase 2: lanes over frames (contiguous floats of one planar row); thread: channels warp + {0, 8, 16, 24}, 4 x 32 frames
  float o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t cl = warp + 8 * (i >> 2), c = c0 + cl, fl = (i & 3) * 32 + lane;
    o[i] = 0.f;
    if (c < a.n_inputs) o[i] = is_prev ? a.xin_prev[(uint64_t)c * a.xstride + a.prev_off + f0 + fl] : tile[fl][cl];
  }
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t cl = warp + 8 * (i >> 2), c = c0 + cl, fl = (i & 3) * 32 + lane;
    if (c < a.n_inputs) a.xin_cur[(uint64_t)c * a.xstride + f0 + fl] = o[i];
  }
}

// ------------------------------------------------------------------------------------------
// k_rfft : windows of 2B floats -> packed spectra
// ------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB)
k_rfft(const float* __restrict__ src, uint64_t ch_stride, uint32_t win_stride, float2* __restrict__ dst, uint64_t dst_ch_stride,
       uint32_t R, uint32_t slot0, const float2* __restrict__ tw, float scale, uint32_t nch) {
  constexpr int RAD = FftCfg<M>::R, NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP;
  __shared__ float2 smem[FPB][MP];
  float2* s = smem[threadIdx.y];
  const int tid = threadIdx.x;
  const uint32_t chq = blockIdx.x * FPB + threadIdx.y, t = blockIdx.y;
  const bool active = chq < nch;
  const uint32_t ch = active ? chq : nch - 1;  // idle transforms of the last CTA recompute a valid one, stores masked
  const float2* win = reinterpret_cast<const float2*>(src + ch * ch_stride + (uint64_t)t * win_stride);
#pragma unroll
  for (int h = 0; h < RAD; h++) s[PAD(tid + h * NT)] = win[tid + h * NT];
  __syncthreads();
  cfft_smem<M, false>(s, tw, tid);
  const uint32_t slot = (slot0 + t) % R;
  rfft_split_store<M>(s, tw, dst + ch * dst_ch_stride + (uint64_t)slot * M, scale, tid, active);
}

// Radix-8 sizes: persistent CTAs loop over (channel group, block) items with their twiddles in registers, the next
// item's window prefetched into registers, the first pass straight from those registers, and (for M = 512) one named
// barrier per transform instead of the block barrier.
template <int M>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB)
k_rfft8(const float* __restrict__ src, uint64_t ch_stride, uint32_t win_stride, float2* __restrict__ dst, uint64_t dst_ch_stride,
        uint32_t R, uint32_t slot0, const float2* __restrict__ tw, float scale, uint32_t nch, uint32_t T) {
  constexpr int NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP;
  __shared__ float2 smem[FPB][MP];
  float2* s = smem[threadIdx.y];
  const int tid = threadIdx.x;
  Tw8<M> tw8;
  load_tw8<M>(tw8, tw, tid);
  const uint32_t nchg = ceil_div_dev(nch, (uint32_t)FPB), nitems = nchg * T;
  auto window = [&](uint32_t item, uint32_t& ch, uint32_t& t, bool& active) {
    t = item / nchg;
    const uint32_t chq = (item - t * nchg) * FPB + threadIdx.y;
    active = chq < nch;
    ch = active ? chq : nch - 1;  // idle transforms of a last group recompute a valid one, stores masked
    return reinterpret_cast<const float2*>(src + ch * ch_stride + (uint64_t)t * win_stride);
  };
  uint32_t item = blockIdx.x, ch = 0, t = 0;
  bool active = false;
  float2 vn[8];
  if (item < nitems) {
    const float2* win = window(item, ch, t, active);
#pragma unroll
    for (int r = 0; r < 8; r++) vn[r] = win[tid + r * NT];
  }
  for (; item < nitems; item += gridDim.x) {
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = vn[r];
    const uint32_t ch_c = ch, t_c = t;
    const bool active_c = active;
    if (item + gridDim.x < nitems) {
      const float2* win = window(item + gridDim.x, ch, t, active);
#pragma unroll
      for (int r = 0; r < 8; r++) vn[r] = win[tid + r * NT];
    }
    fft_bar<M>();  // the previous item's split stage is done with the workspace
    pass8_first<M, false>(v, s, tid);
    passes8_rest<M, false>(s, tw8, tid);
    const uint32_t slot = (slot0 + t_c) % R;
    rfft_split_store8<M>(s, tw8, dst + ch_c * dst_ch_stride + (uint64_t)slot * M, scale, tid, active_c);
  }
}

// ------------------------------------------------------------------------------------------
// k_fdl_mac : the hot kernel
// ------------------------------------------------------------------------------------------
struct MacSeg {
  const float4* H;    // filter spectra, row 0 (row stride = B/2 float4)
  uint32_t fdl_ch;    // input channel whose FDL this term reads
  uint32_t p0, np;    // partition range of this segment
  uint32_t slot;      // partial-sum slot the run is written to
  uint32_t flags;     // bit0: reset accumulator before, bit1: write accumulator after
  uint32_t pad;
};
static_assert(sizeof(MacSeg) == 32, "MacSeg layout");

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

// One complex multiply-accumulate acc += h * x: four FMAs in a fixed order (every MAC kernel uses exactly this
// sequence per output, so their results are bit-identical):
//   re = fma(hr, xr, re); re = fma(-hi, xi, re); im = fma(hi, xr, im); im = fma(hr, xi, im)
//
// Bin 0 of a packed row holds (DC, Nyquist), two REAL spectra.  The MAC kernels do not special-case it: column 0
// runs the generic complex MAC, whose real part is G = sum DCh DCx - sum Nqh Nqx, and the Nyquist sum
// N = sum Nqh Nqx is accumulated separately (one extra FMA per row in the streaming kernel, k_nyq_mac next to the
// time-batched kernel; same order, same fma).  k_irfft restores bin 0 = (G + N, N).  This keeps selects and
// register-pair shuffles out of the hot loops (profiles/: ALU pipe 45 % -> see DESIGN.md).
__device__ __forceinline__ void cmac(float& re, float& im, float hr, float hi, float xr, float xi) {
  re = fmaf(hr, xr, re);
  re = fmaf(-hi, xi, re);
  im = fmaf(hi, xr, im);
  im = fmaf(hr, xi, im);
}

// The same complex MAC as two packed FP32x2 FMAs (Blackwell FFMA2: one instruction, two lanes):
//   (re, im) += (hr, hi) * xr ;  (re, im) += (-hi, hr) * xi
// ptxas folds the scalar broadcast and the swapped / negated pair into FFMA2 operand modifiers, so no extra
// registers or moves are needed.  Each lane is an IEEE fma: bit-identical to cmac().
__device__ __forceinline__ void cmac_x2(float2& acc, float2 h, float2 x) {
  acc = __ffma2_rn(h, make_float2(x.x, x.x), acc);
  acc = __ffma2_rn(make_float2(-h.y, h.x), make_float2(x.y, x.y), acc);
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

// ------------------------------------------------------------------------------------------
// k_irfft : partial sums -> time domain -> crossfade -> delay ring
// ------------------------------------------------------------------------------------------
struct PlanView {
  const uint32_t* job_slot_first;
  const uint32_t* job_slot_count;
  const uint32_t* xjob;  // per stream: extra job to crossfade into, kNoJob, or kSameJob
};

template <int M>
__device__ __forceinline__ void job_to_block(const float2* __restrict__ ypart_t, const float* __restrict__ nyq_t, uint64_t s_stride,
                                             uint32_t first, uint32_t count, float2* __restrict__ x,
                                             float2* __restrict__ s, const float2* __restrict__ tw, int tid,
                                             float (&o)[FftCfg<M>::R]) {
  constexpr int RAD = FftCfg<M>::R, NT = FftCfg<M>::NT;
#pragma unroll
  for (int h = 0; h < RAD; h++) {
    const int k = tid + h * NT;
    float2 a = make_float2(0.f, 0.f);
    // fixed slot order (deterministic sums); four loads in flight per step
    for (uint32_t sl = 0; sl < count; sl += 4) {
      float2 v[4];
#pragma unroll
      for (int q = 0; q < 4; q++)
        v[q] = (sl + q < count) ? ypart_t[(uint64_t)(first + sl + q) * s_stride + k] : make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (sl + q < count) {
          a.x += v[q].x;
          a.y += v[q].y;
        }
    }
    if (k == 0 && nyq_t) {
      // bin 0: the MAC kernels left G = DC - N in the real part; add the Nyquist sum back (same slot order)
      // (nyq_t == NULL: the tensor-core MIMO path writes bin 0 = (DC, Nyquist) directly)
      float n = 0.f;
      for (uint32_t sl = 0; sl < count; sl++) n += nyq_t[first + sl];
      a = make_float2(a.x + n, n);
    }
    x[k] = a;
  }
  __syncthreads();
  irfft_unsplit<M>(x, tw, s, tid);
  __syncthreads();
  cfft_smem<M, true>(s, tw, tid);
  // overlap-save: y[B+n] = component (n&1) of z[M/2 + n/2]; this thread keeps n = 2(tid + h NT) + {0,1}, h < R/2
#pragma unroll
  for (int h = 0; h < RAD / 2; h++) {
    float2 z = s[PAD(M / 2 + tid + h * NT)];
    o[2 * h] = z.x;
    o[2 * h + 1] = z.y;
  }
  __syncthreads();
}

template <int M>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB)
k_irfft(const float2* __restrict__ ypart, uint32_t slot_stride, PlanView first_blk, PlanView steady, uint32_t n_first,
        const float2* __restrict__ tw, float* __restrict__ ybuf, uint32_t Rd, uint32_t wpos0, uint32_t n_streams,
        const float* __restrict__ nyq_part, uint64_t t_stride, uint64_t s_stride, uint32_t stream0) {
  // spectrum of (block t, slot s) at ypart + t * t_stride + s * s_stride (the MAC kernels: t_stride = slot_stride * M,
  // s_stride = M; after the sharded reduce-scatter: [slot][t][M]); streams stream0 .. stream0 + n_streams - 1
  constexpr int RAD = FftCfg<M>::R, NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP;
  extern __shared__ float2 k_irfft_smem[];  // per transform: summed spectrum x[M] + padded FFT workspace s[MP]
  float2* x = k_irfft_smem + (size_t)threadIdx.y * (M + MP);
  float2* s = x + M;
  const int tid = threadIdx.x;
  const uint32_t sq = blockIdx.x * FPB + threadIdx.y, t = blockIdx.y;
  const bool active = sq < n_streams;
  const uint32_t stream = stream0 + (active ? sq : n_streams - 1);
  const bool first = t < n_first;  // blocks covered by the transitional plan (0 or 1 of them)
  const PlanView pv = first ? first_blk : steady;
  const float2* ypart_t = ypart + (uint64_t)t * t_stride;
  const float* nyq_t = nyq_part ? nyq_part + (uint64_t)t * slot_stride : nullptr;
  float o[RAD];
  job_to_block<M>(ypart_t, nyq_t, s_stride, pv.job_slot_first[stream], pv.job_slot_count[stream], x, s, tw, tid, o);
  // the crossfade decision must be uniform across the CTA (block-wide barriers inside job_to_block):
  // every transform of the CTA runs the second pass when any of them needs it
  const uint32_t xj = first ? pv.xjob[stream] : kNoJob;
  const int any_x = __syncthreads_or(xj != kNoJob && xj != kSameJob);
  float o2[RAD];
  if (any_x) {
    const bool mine = (xj != kNoJob && xj != kSameJob);
    job_to_block<M>(ypart_t, nyq_t, s_stride, mine ? pv.job_slot_first[xj] : 0u, mine ? pv.job_slot_count[xj] : 0u, x, s, tw, tid, o2);
  }
  if (xj != kNoJob) {
    if (xj == kSameJob) {
#pragma unroll
      for (int i = 0; i < RAD; i++) o2[i] = o[i];
    }
    // out = (1-g) o_f + g o_f', g_n = n/B  (MixSamples + Interpolator ramp, sampled before the step)
    const float inc = 1.0f / (float)M;
#pragma unroll
    for (int h = 0; h < RAD / 2; h++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const uint32_t n = 2 * (tid + h * NT) + c;
        const float g = __fmul_rn((float)n, inc);
        const float a = __fmul_rn(__fsub_rn(1.0f, g), o[2 * h + c]);
        const float b = __fmul_rn(g, o2[2 * h + c]);
        o[2 * h + c] = __fadd_rn(a, b);
      }
  }
  if (!active) return;
  float* ring = ybuf + (uint64_t)stream * Rd;
  const uint32_t w = (wpos0 + t * (uint32_t)M) % Rd;
#pragma unroll
  for (int h = 0; h < RAD / 2; h++)
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const uint32_t n = 2 * (tid + h * NT) + c;
      uint32_t idx = w + n;
      if (idx >= Rd) idx -= Rd;
      ring[idx] = o[2 * h + c];
    }
}

// Radix-8 sizes: same contract as job_to_block, with cached twiddles, all of a slot's partial loads in flight at once,
// the inverse split and the first pass in registers, per-transform barriers.
template <int M>
__device__ __forceinline__ void job_to_block8(const float2* __restrict__ ypart_t, const float* __restrict__ nyq_t, uint64_t s_stride,
                                              uint32_t first, uint32_t count, float2* __restrict__ x, float2* __restrict__ s,
                                              const Tw8<M>& tw8, int tid, float (&o)[8]) {
  constexpr int NT = M / 8;
  float2 a[8];
#pragma unroll
  for (int h = 0; h < 8; h++) a[h] = make_float2(0.f, 0.f);
  for (uint32_t sl = 0; sl < count; sl++) {  // fixed slot order (deterministic sums)
    const float2* row = ypart_t + (uint64_t)(first + sl) * s_stride;
    float2 v[8];
#pragma unroll
    for (int h = 0; h < 8; h++) v[h] = row[tid + h * NT];
#pragma unroll
    for (int h = 0; h < 8; h++) {
      a[h].x += v[h].x;
      a[h].y += v[h].y;
    }
  }
  if (tid == 0 && nyq_t) {
    // bin 0: the MAC kernels left G = DC - N in the real part; add the Nyquist sum back (same slot order)
    float n = 0.f;
    for (uint32_t sl = 0; sl < count; sl++) n += nyq_t[first + sl];
    a[0] = make_float2(a[0].x + n, n);
  }
#pragma unroll
  for (int h = 0; h < 8; h++) x[tid + h * NT] = a[h];
  fft_bar<M>();  // x complete; also: every thread of the transform is past its reads of s from the previous job
  float2 v[8];
  irfft_unsplit8<M>(x, tw8, v, tid);
  pass8_first<M, true>(v, s, tid);
  passes8_rest<M, true>(s, tw8, tid);
  // overlap-save: y[B+n] = component (n&1) of z[M/2 + n/2]; this thread keeps n = 2(tid + h NT) + {0,1}, h < 4
#pragma unroll
  for (int h = 0; h < 4; h++) {
    float2 z = s[PAD(M / 2 + tid + h * NT)];
    o[2 * h] = z.x;
    o[2 * h + 1] = z.y;
  }
}

template <int M>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB)
k_irfft8(const float2* __restrict__ ypart, uint32_t slot_stride, PlanView first_blk, PlanView steady, uint32_t n_first,
         const float2* __restrict__ tw, float* __restrict__ ybuf, uint32_t Rd, uint32_t wpos0, uint32_t n_streams,
         const float* __restrict__ nyq_part, uint64_t t_stride, uint64_t s_stride, uint32_t stream0, uint32_t T) {
  constexpr int RAD = 8, NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP;
  extern __shared__ float2 k_irfft_smem[];  // per transform: summed spectrum x[M] + padded FFT workspace s[MP]
  float2* x = k_irfft_smem + (size_t)threadIdx.y * (M + MP);
  float2* s = x + M;
  const int tid = threadIdx.x;
  Tw8<M> tw8;
  load_tw8<M>(tw8, tw, tid);
  const uint32_t nsg = ceil_div_dev(n_streams, (uint32_t)FPB), nitems = nsg * T;
  for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const uint32_t t = item / nsg, sq = (item - t * nsg) * FPB + threadIdx.y;
    const bool active = sq < n_streams;
    const uint32_t stream = stream0 + (active ? sq : n_streams - 1);
    const bool first = t < n_first;  // blocks covered by the transitional plan (0 or 1 of them)
    const PlanView pv = first ? first_blk : steady;
    const float2* ypart_t = ypart + (uint64_t)t * t_stride;
    const float* nyq_t = nyq_part ? nyq_part + (uint64_t)t * slot_stride : nullptr;
    float o[RAD];
    job_to_block8<M>(ypart_t, nyq_t, s_stride, pv.job_slot_first[stream], pv.job_slot_count[stream], x, s, tw8, tid, o);
    const uint32_t xj = first ? pv.xjob[stream] : kNoJob;
    // the barriers inside job_to_block8 span one transform (M = 512) or the CTA: the decision to run the second
    // pass is made CTA-uniform so that both cases are safe
    const int any_x = __syncthreads_or(xj != kNoJob && xj != kSameJob);
    float o2[RAD];
    if (any_x) {
      const bool mine = (xj != kNoJob && xj != kSameJob);
      job_to_block8<M>(ypart_t, nyq_t, s_stride, mine ? pv.job_slot_first[xj] : 0u, mine ? pv.job_slot_count[xj] : 0u, x, s, tw8,
                       tid, o2);
    }
    if (xj != kNoJob) {
      if (xj == kSameJob) {
#pragma unroll
        for (int i = 0; i < RAD; i++) o2[i] = o[i];
      }
      // out = (1-g) o_f + g o_f', g_n = n/B  (MixSamples + Interpolator ramp, sampled before the step)
      const float inc = 1.0f / (float)M;
#pragma unroll
      for (int h = 0; h < RAD / 2; h++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const uint32_t n = 2 * (tid + h * NT) + c;
          const float g = __fmul_rn((float)n, inc);
          const float a = __fmul_rn(__fsub_rn(1.0f, g), o[2 * h + c]);
          const float b = __fmul_rn(g, o2[2 * h + c]);
          o[2 * h + c] = __fadd_rn(a, b);
        }
    }
    if (active) {
      float* ring = ybuf + (uint64_t)stream * Rd;
      const uint32_t w = (wpos0 + t * (uint32_t)M) % Rd;
#pragma unroll
      for (int h = 0; h < RAD / 2; h++) {
        const uint32_t n = 2 * (tid + h * NT);
        uint32_t idx = w + n;  // w and n are even, Rd is a multiple of the block size: the pair never straddles the wrap
        if (idx >= Rd) idx -= Rd;
        if ((Rd & 1u) == 0 && (w & 1u) == 0) {
          *reinterpret_cast<float2*>(ring + idx) = make_float2(o[2 * h], o[2 * h + 1]);
        } else {
          ring[idx] = o[2 * h];
          uint32_t i1 = idx + 1;
          if (i1 >= Rd) i1 -= Rd;
          ring[i1] = o[2 * h + 1];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_pcm_out : delay read + mixdown + format conversion
// ------------------------------------------------------------------------------------------
// one routed path in mixdown order (ascending stream per output == MixSamples call order), everything k_pcm_out
// needs about it in one place
struct RouteEntry {
  uint32_t stream;  // delay ring of the path
  float gain;
  uint32_t icur, iold;  // (uint32)delay mod Rd for the integer-delay mode: in force after / before this call's first block boundary
  uint32_t flags;       // bit0: crossfade old->cur over the first block
  uint32_t pad;
  double dcur, dold;    // the same delays in samples (fractional mode)
};
static_assert(sizeof(RouteEntry) == 40, "RouteEntry layout");

struct RouteView {
  const uint32_t* out_first;  // [n_outputs+1] CSR over outputs
  const RouteEntry* entry;    // per route
};

struct PcmOutArgs {
  uint8_t* pcm;
  int fmt;
  int be;
  uint32_t out_channels, n_outputs;
  uint32_t B, T;
  const float* ybuf;
  uint32_t Rd, wpos0;
  int fractional;
  int fast;  // typed stores (see PcmInArgs::fast)
  RouteView rv;
};

__device__ __forceinline__ float delayed_read(const float* __restrict__ ring, uint32_t Rd, uint32_t w, uint32_t n, double d,
                                              uint32_t di, int fractional) {
  if (fractional) {
    // FractionalSample(ring, 0, 1, Rd, fmod((w + n + Rd) - d, Rd))   (src/FractionalSample.cpp:312-341)
    const double pos = fmod((double)(w + n + Rd) - d, (double)Rd);
    return __double2float_rn(fractional_sample_dev<float>(ring, 0, 1, Rd, pos));
  }
  // ring[(w + n - d) mod R] with d = (uint)delay mod R precomputed on the host   (src/SoundDelayBuffer.cpp:141)
  uint32_t idx = w + n + Rd - di;  // w < Rd, n < B <= Rd, di < Rd  ->  idx < 3 Rd
  if (idx >= Rd) idx -= Rd;
  if (idx >= Rd) idx -= Rd;
  return ring[idx];
}

static constexpr uint32_t kPcmOutCache = 64;  // routes of one 32-output tile kept in shared memory

template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_out(PcmOutArgs a) {
  __shared__ float tile[32][33];
  __shared__ uint32_t s_first[33];
  __shared__ RouteEntry s_rt[kPcmOutCache];
  const uint32_t f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t t = f0 / a.B;               // a tile lies inside one block (B % 32 == 0)
  const uint32_t w = (a.wpos0 + t * a.B) % a.Rd;
  const float inc = 1.0f / (float)a.B;
  // the tile's slice of the route tables -> shared memory (two dependent loads per CTA instead of a chain of table
  // lookups per sample); tiles with more than kPcmOutCache routes read the entries from global memory
  const uint32_t no = min(32u, a.n_outputs - c0);
  if (threadIdx.x <= no) s_first[threadIdx.x] = a.rv.out_first[c0 + threadIdx.x];
  __syncthreads();
  const uint32_t r0 = s_first[0], nr = s_first[no] - r0;
  const bool cached = nr <= kPcmOutCache;
  if (cached && threadIdx.x < nr) s_rt[threadIdx.x] = a.rv.entry[r0 + threadIdx.x];
  __syncthreads();
  // phase 1: lanes over frames (ring reads are contiguous), one output channel per warp pass
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    const uint32_t cl = warp + 8 * i, o = c0 + cl;
    float bus = 0.f;
    if (o < a.n_outputs) {
      const uint32_t n = f0 + lane - t * a.B;  // frame inside the block
      const uint32_t rb = s_first[cl], re = s_first[cl + 1];
      for (uint32_t r = rb; r < re; r++) {  // ascending stream order == MixSamples call order
        const RouteEntry en = cached ? s_rt[r - r0] : a.rv.entry[r];
        if (!(en.gain != 0.0f)) continue;  // (mul != T()): a zero gain is a no-op (src/SoundMixing.h:65-69)
        const float* ring = a.ybuf + (uint64_t)en.stream * a.Rd;
        float v = delayed_read(ring, a.Rd, w, n, en.dcur, en.icur, a.fractional);
        if (t == 0 && (en.flags & 1u)) {
          const float vo = delayed_read(ring, a.Rd, w, n, en.dold, en.iold, a.fractional);
          const float g = __fmul_rn((float)n, inc);
          v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, g), vo), __fmul_rn(g, v));
        }
        bus = __fadd_rn(bus, __fmul_rn(en.gain, v));  // dst += mul * src, rounded separately
      }
    }
    tile[lane][cl] = bus;
  }
  __syncthreads();
  // phase 2: lanes over channels (contiguous bytes of one interleaved frame)
  constexpr uint32_t bps = FmtBytes<FMT>::value;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t fl = warp + 8 * i, o = c0 + lane;
    if (o >= a.n_outputs) continue;
    const uint32_t frame = f0 + fl;
    store_from_f32_t<FMT, ACC>(a.pcm + ((uint64_t)frame * a.out_channels + o) * bps, tile[fl][lane]);
  }
}

// Mixdown of many paths into few outputs (the binaural renderer: 64 sources x 2 ears -> 2 outputs).  The kernel above
// walks the routes of an output one after the other inside one thread: 64 dependent table + ring reads per sample, and
// only n_outputs of its 32 channel slots do anything.  Here every thread of the CTA takes (route, frame) items: the
// delayed reads, the delay crossfade and the products gain * v of ALL routes of a 32-frame tile are formed in parallel
// into shared memory, then one thread per (output, frame) adds the products in ascending route order -- the same
// dst += mul * src with separately rounded product and sum (src/SoundMixing.h:76-79), zero gains skipped, so the bytes are
// those of k_pcm_out (tests: routed engines run both kernels' shapes against the oracle and each other).
static constexpr uint32_t kMixMaxRoutes = 256;

template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_out_mix(PcmOutArgs a) {
  __shared__ float prod[kMixMaxRoutes][32];
  __shared__ RouteEntry s_rt[kMixMaxRoutes];
  __shared__ uint32_t s_first[33];
  const uint32_t f0 = blockIdx.x * 32;
  const uint32_t t = f0 / a.B;  // a tile lies inside one block (B % 32 == 0)
  const uint32_t w = (a.wpos0 + t * a.B) % a.Rd;
  const float inc = 1.0f / (float)a.B;
  const uint32_t no = a.n_outputs;  // <= 32 (host)
  if (threadIdx.x <= no) s_first[threadIdx.x] = a.rv.out_first[threadIdx.x];
  __syncthreads();
  const uint32_t r0 = s_first[0], nr = s_first[no] - r0;  // <= kMixMaxRoutes (host)
  for (uint32_t r = threadIdx.x; r < nr; r += 256) s_rt[r] = a.rv.entry[r0 + r];
  __syncthreads();
  const uint32_t nb0 = f0 - t * a.B;  // frame of the tile's first sample inside its block
  // stage 1: items (route, frame), four per thread and pass so that their ring reads are in flight together
  for (uint32_t base = threadIdx.x; base < nr * 32; base += 4 * 256) {
    float p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t idx = base + 256 * k;
      p[k] = 0.f;
      if (idx < nr * 32) {
        const uint32_t r = idx >> 5, n = nb0 + (idx & 31);
        const RouteEntry en = s_rt[r];
        if (en.gain != 0.0f) {
          const float* ring = a.ybuf + (uint64_t)en.stream * a.Rd;
          float v = delayed_read(ring, a.Rd, w, n, en.dcur, en.icur, a.fractional);
          if (t == 0 && (en.flags & 1u)) {
            const float vo = delayed_read(ring, a.Rd, w, n, en.dold, en.iold, a.fractional);
            const float g = __fmul_rn((float)n, inc);
            v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, g), vo), __fmul_rn(g, v));
          }
          p[k] = __fmul_rn(en.gain, v);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t idx = base + 256 * k;
      if (idx < nr * 32) prod[idx >> 5][idx & 31] = p[k];
    }
  }
  __syncthreads();
  // stage 2: one thread per (output, frame): the ordered sum, then the sample in the output format
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  for (uint32_t item = threadIdx.x; item < no * 32; item += 256) {
    const uint32_t o = item >> 5, fl = item & 31;
    float bus = 0.f;
    for (uint32_t r = s_first[o] - r0; r < s_first[o + 1] - r0; r++)
      if (s_rt[r].gain != 0.0f) bus = __fadd_rn(bus, prod[r][fl]);  // a zero gain is a no-op, not "+ 0"
    store_from_f32_t<FMT, ACC>(a.pcm + ((uint64_t)(f0 + fl) * a.out_channels + o) * bps, bus);
  }
}

// 128-frame tiles (B % 128 == 0).  Outputs fed by exactly one path with an integer delay and no delay crossfade in this
// block (every output of the PER_CHANNEL and MIMO modes in the steady state) issue their four ring reads together; the
// arithmetic is the same dst += mul * src, rounded separately.
template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_out128(PcmOutArgs a) {
  __shared__ float tile[128][33];
  __shared__ uint32_t s_first[33];
  __shared__ RouteEntry s_rt[kPcmOutCache];
  const uint32_t f0 = blockIdx.x * 128, c0 = blockIdx.y * 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t t = f0 / a.B;               // a tile lies inside one block (B % 128 == 0)
  const uint32_t w = (a.wpos0 + t * a.B) % a.Rd;
  const float inc = 1.0f / (float)a.B;
  const uint32_t no = min(32u, a.n_outputs - c0);
  if (threadIdx.x <= no) s_first[threadIdx.x] = a.rv.out_first[c0 + threadIdx.x];
  __syncthreads();
  const uint32_t r0 = s_first[0], nr = s_first[no] - r0;
  const bool cached = nr <= kPcmOutCache;
  if (cached && threadIdx.x < nr) s_rt[threadIdx.x] = a.rv.entry[r0 + threadIdx.x];
  __syncthreads();
  // phase 1: lanes over frames; thread: outputs warp + {0, 8, 16, 24}, 4 x 32 frames each
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t cl = warp + 8 * q, o = c0 + cl;
    float bus[4] = {0.f, 0.f, 0.f, 0.f};
    if (o < a.n_outputs) {
      const uint32_t rb = s_first[cl], re = s_first[cl + 1];
      const uint32_t nb = f0 - t * a.B + lane;  // frame inside the block of the first of the four chunks
      bool done = false;
      if (re == rb + 1 && !a.fractional) {
        const RouteEntry en = cached ? s_rt[rb - r0] : a.rv.entry[rb];
        if (!(t == 0 && (en.flags & 1u))) {
          done = true;
          if (en.gain != 0.0f) {
            const float* ring = a.ybuf + (uint64_t)en.stream * a.Rd;
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = delayed_read(ring, a.Rd, w, nb + 32 * k, 0.0, en.icur, 0);
#pragma unroll
            for (int k = 0; k < 4; k++) bus[k] = __fadd_rn(0.f, __fmul_rn(en.gain, v[k]));
          }
        }
      }
      if (!done) {
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
          const uint32_t n = nb + 32 * k;
          float b = 0.f;
          for (uint32_t r = rb; r < re; r++) {  // ascending stream order == MixSamples call order
            const RouteEntry en = cached ? s_rt[r - r0] : a.rv.entry[r];
            if (!(en.gain != 0.0f)) continue;  // (mul != T()): a zero gain is a no-op (src/SoundMixing.h:65-69)
            const float* ring = a.ybuf + (uint64_t)en.stream * a.Rd;
            float v = delayed_read(ring, a.Rd, w, n, en.dcur, en.icur, a.fractional);
            if (t == 0 && (en.flags & 1u)) {
              const float vo = delayed_read(ring, a.Rd, w, n, en.dold, en.iold, a.fractional);
              const float g = __fmul_rn((float)n, inc);
              v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, g), vo), __fmul_rn(g, v));
            }
            b = __fadd_rn(b, __fmul_rn(en.gain, v));  // dst += mul * src, rounded separately
          }
          bus[k] = b;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) tile[32 * k + lane][cl] = bus[k];
  }
  __syncthreads();
  // phase 2: lanes over channels (contiguous bytes of one interleaved frame)
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  const uint32_t o = c0 + lane;
  if (o < a.n_outputs) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t fl = warp + 8 * i;
      store_from_f32_t<FMT, ACC>(a.pcm + ((uint64_t)(f0 + fl) * a.out_channels + o) * bps, tile[fl][lane]);
    }
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

__global__ void k_flush(float4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

// Input-sharded MIMO: this rank's partial output spectra, job by job (slots summed in fixed order, bin 0 restored to
// (DC, Nyquist)), into the reduce-scatter send buffer [output][t][B] -- one contiguous chunk per destination rank.
__global__ void __launch_bounds__(256) k_gather_spectra(const float2* __restrict__ ypart, const float* __restrict__ nyq_part,
                                                        uint32_t slot_stride, PlanView pv, float2* __restrict__ send, uint32_t B,
                                                        uint32_t T) {
  const uint32_t o = blockIdx.x, t = blockIdx.y;
  const uint32_t first = pv.job_slot_first[o], count = pv.job_slot_count[o];
  const float2* yt = ypart + (uint64_t)t * slot_stride * B;
  for (uint32_t k = threadIdx.x; k < B; k += blockDim.x) {
    float2 a = make_float2(0.f, 0.f);
    for (uint32_t sl = 0; sl < count; sl++) {
      const float2 v = yt[(uint64_t)(first + sl) * B + k];
      a.x += v.x;
      a.y += v.y;
    }
    if (k == 0 && nyq_part) {
      float n = 0.f;
      for (uint32_t sl = 0; sl < count; sl++) n += nyq_part[(uint64_t)t * slot_stride + first + sl];
      a = make_float2(a.x + n, n);
    }
    send[((uint64_t)o * T + t) * B + k] = a;
  }
}

// ---- peer-memory mixdown: the reduce of the input-sharded MIMO engine without a collective library ----------------------
// Every rank adds its partial slots like k_gather_spectra, but stores the spectrum of output o straight into the memory of
// the rank that owns o (NVLink peer stores into a buffer opened with cudaIpcOpenMemHandle), at slot (o_local, source rank).
// The owner's inverse-transform kernel then adds the `world` slots of an output in rank order -- the same fixed-order slot
// sum it already runs over the MAC's partial sums, so the result does not depend on a collective's reduction schedule.
// Completion: the last CTA of a launch publishes the call's epoch in every peer's flag array after a system-scope fence;
// k_peer_wait (one warp, in stream order before the inverse transforms) spins until all sources have published it.  The
// receive buffer is double-buffered by epoch parity: a source can only be two calls ahead after it has seen this rank's
// flag of the call in between, which this rank publishes after its own inverse transforms of the older call (stream order).
struct PeerTable {
  float2* data[16];     // receive buffers of the ranks (own rank: the local buffer)
  uint32_t* flags[16];  // their flag arrays, [2][world]
};

__global__ void __launch_bounds__(256) k_gather_spectra_peer(const float2* __restrict__ ypart, const float* __restrict__ nyq_part,
                                                             uint32_t slot_stride, PlanView pv, PeerTable pt, uint32_t world,
                                                             uint32_t rank, uint32_t nloc, uint32_t B, uint32_t T, uint64_t half,
                                                             uint32_t parity, uint32_t epoch, uint32_t* __restrict__ done) {
  const uint32_t o = blockIdx.x, t = blockIdx.y;
  const uint32_t first = pv.job_slot_first[o], count = pv.job_slot_count[o];
  const float2* yt = ypart + (uint64_t)t * slot_stride * B;
  const uint32_t r = o / nloc, ol = o - r * nloc;
  float2* dst = pt.data[r] + (uint64_t)parity * half + (((uint64_t)ol * world + rank) * T + t) * B;
  for (uint32_t k = threadIdx.x; k < B; k += blockDim.x) {
    float2 a = make_float2(0.f, 0.f);
    for (uint32_t sl = 0; sl < count; sl++) {
      const float2 v = yt[(uint64_t)(first + sl) * B + k];
      a.x += v.x;
      a.y += v.y;
    }
    if (k == 0 && nyq_part) {
      float n = 0.f;
      for (uint32_t sl = 0; sl < count; sl++) n += nyq_part[(uint64_t)t * slot_stride + first + sl];
      a = make_float2(a.x + n, n);
    }
    dst[k] = a;
  }
  __threadfence_system();  // this thread's peer stores are performed before the CTA counts itself done
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t total = gridDim.x * gridDim.y;
    if (atomicAdd(done, 1u) == total - 1) {
      *done = 0;  // ready for the next launch (stream-ordered)
      __threadfence_system();
      for (uint32_t q = 0; q < world; q++) *((volatile uint32_t*)pt.flags[q] + parity * world + rank) = epoch;
    }
  }
}

__global__ void __launch_bounds__(32) k_peer_wait(const uint32_t* flags, uint32_t world, uint32_t parity, uint32_t epoch,
                                                  int* status) {
  if (threadIdx.x < world) {
    const volatile uint32_t* f = flags + parity * world + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(*f - epoch) < 0) {
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a source never arrived; bbx_engine_sync reports it
        *status = 1 + (int)threadIdx.x;
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

// comm.cu
int comm_reduce_scatter_f32(bbx_comm* c, const float* send, float* recv, size_t recvcount, cudaStream_t st);
int comm_world(const bbx_comm* c);
int comm_rank(const bbx_comm* c);

}  // namespace bbx

using namespace bbx;

// ==========================================================================================
// host side
// ==========================================================================================
struct bbx_filter {
  bbx_engine* engine;
  float2* H;  // device [P][B]
  uint32_t P;
};

struct PathState {
  uint32_t input = 0, output = 0;
  float gain = 1.0f;
  const bbx_filter* cur = nullptr;
  const bbx_filter* pend = nullptr;
  bool has_pending = false, xfade = false;
  double delay = 0.0, pend_delay = 0.0;
};

// one job = one accumulated output spectrum: a list of (filter, input) terms
struct JobTerm {
  const bbx_filter* f;
  uint32_t input;
};

// Ring of pinned upload buffers for one device table.  Every upload is a copy on the engine stream followed by an event;
// re-using a slot waits only for the copy that last read it (kDepth uploads ago), so the host can prepare the plans and
// routes of the next calls while the device still works on the previous ones (dynamic IR switching: one upload set per
// switch).
struct Staging {
  static constexpr int kDepth = 4;
  uint8_t* h[kDepth] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[kDepth] = {nullptr, nullptr, nullptr, nullptr};
  bool pending[kDepth] = {false, false, false, false};
  int cur = 0;
  size_t bytes = 0;
};

// device-resident MAC plan + host mirror
struct MacPlan {
  // host staging (pinned ring) and device blob, same layout
  Staging stg;
  uint8_t* h_blob = nullptr;  // the slot being filled by build_plan
  uint8_t* d_blob = nullptr;
  size_t blob_bytes = 0;
  // offsets inside the blob
  size_t off_segs = 0, off_cta = 0, off_first = 0, off_count = 0, off_xjob = 0;
  uint32_t n_ctas = 0, n_slots = 0, n_jobs = 0, total_rows = 0, n_terms = 0;
  uint32_t occ = 1;  // streaming-MAC variant this plan was cut for (CTAs per SM <-> unroll depth)
  bool valid = false;
  const MacSeg* segs() const { return (const MacSeg*)(d_blob + off_segs); }
  const uint32_t* cta_seg_begin() const { return (const uint32_t*)(d_blob + off_cta); }
  PlanView view() const {
    PlanView v;
    v.job_slot_first = (const uint32_t*)(d_blob + off_first);
    v.job_slot_count = (const uint32_t*)(d_blob + off_count);
    v.xjob = (const uint32_t*)(d_blob + off_xjob);
    return v;
  }
};

struct bbx_engine {
  bbx_config cfg;
  int device = 0;
  uint32_t B = 0, Pmax = 0, n_in = 0, n_out = 0, n_paths = 0, n_streams = 0, Tmax = 1;
  uint32_t R = 0;   // FDL ring slots
  uint32_t Rd = 0;  // delay ring frames
  uint32_t xstride = 0;
  int mode = 0;
  cudaStream_t stream = nullptr;
  // device buffers
  float2* tw = nullptr;
  float* xin[2] = {nullptr, nullptr};
  float2* fdl = nullptr;
  float2* ypart = nullptr;
  float* nyq_part = nullptr;  // [Tmax][max_slots] Nyquist partial sums (see cmac)
  float* ybuf = nullptr;
  // host-pointer path: double-buffered PCM staging, copies on their own streams so that the H2D of call
  // n+1 and the D2H of call n-1 overlap the kernels of call n
  uint8_t* d_in[2] = {nullptr, nullptr};
  uint8_t* d_out[2] = {nullptr, nullptr};
  size_t d_io_bytes = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaStream_t s_aux = nullptr;  // side stream: k_nyq_mac runs next to the time-batched MAC
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  cudaEvent_t ev_join_in = nullptr, ev_join_out = nullptr;
  uint64_t host_calls = 0;
  // short host calls: the PCM kernels read / write the caller's pinned (device-mapped) buffers directly over PCIe
  // instead of going through the copy engines and the staging buffers; 0 disables
  size_t direct_io_max_bytes = 1u << 20;
  uint64_t direct_calls = 0;
  float4* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // route tables (device blob + pinned staging)
  Staging route_stg;
  uint32_t n_routes_pcm = 0;  // routes feeding the PCM outputs (set by upload_routes)
  bool pcm_out_mix = true;  // bbx_engine_set_mixdown_kernel: false keeps mixdowns on the per-output kernel
  uint8_t* h_route = nullptr;  // the slot being filled by upload_routes
  uint8_t* d_route = nullptr;
  size_t route_bytes = 0, roff_first = 0, roff_stream = 0, roff_gain = 0, roff_dcur = 0, roff_dold = 0, roff_flags = 0,
         roff_icur = 0, roff_iold = 0, roff_entry = 0;
  bool route_dirty = true;
  // plans
  MacPlan plan_first, plan_steady;
  uint32_t max_segs = 0, max_slots = 0, max_ctas = 0, max_jobs = 0;
  bool steady_dirty = true;
  // state
  std::vector<PathState> paths;
  uint32_t head = 0, wpos = 0, parity = 0, tprev = 1;
  // measurement
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  cudaEvent_t ev_upload = nullptr;  // last H2D copy out of the pinned plan/route staging
  bool upload_pending = false;
  uint32_t mac_occ = 1;          // resident streaming-MAC CTAs per SM; the plan has 148 * mac_occ row ranges
  float mac_l2_keep = 3.f / 16;  // fraction of H / FDL lines given L2 evict-last priority by the streaming MAC
  uint32_t mac_time_tile = 16;   // TT of the time-batched MAC (0 = streaming kernel only)
  uint64_t launches = 0;
  bool profile_mac = false;
  std::vector<cudaEvent_t> mac_events;  // pairs
  size_t mac_events_used = 0;
  double mac_ms_total = 0.0;
  uint64_t mac_launches = 0, mac_units = 0, mac_bytes = 0;
  int last_infmt = FMT_F32, last_outfmt = FMT_F32;
  // input-sharded MIMO (bbx_config::mimo_shard_*): partial spectra -> reduce-scatter -> local outputs only
  uint32_t sh_world = 1, sh_rank = 0, sh_o0 = 0, sh_nloc = 0;  // local outputs [sh_o0, sh_o0 + sh_nloc)
  uint32_t n_out_pcm = 0;                                      // channels written by bbx_process (= sh_nloc when sharded)
  bbx_comm* comm = nullptr;
  // peer-memory mixdown (bbx_engine_peer_export / _attach): receive buffer [2][sh_nloc][world][Tmax][B] + flags [2][world]
  bool px_on = false;
  uint8_t* px_mem = nullptr;       // one allocation: data, then the flags
  size_t px_half = 0;              // float2 elements of one parity half
  size_t px_flag_off = 0;          // byte offset of the flags
  void* px_peer[16] = {nullptr};   // opened peer allocations (own rank: nullptr)
  PeerTable px_table;
  uint32_t px_epoch = 0;
  uint32_t* px_done = nullptr;     // last-CTA counter of k_gather_spectra_peer
  uint32_t* px_view = nullptr;     // device: first[n_out] | count[n_out] | xjob[n_out] for the world slots per local output
  int* px_status_h = nullptr;      // mapped pinned word: k_peer_wait timed out
  int* px_status = nullptr;
  float2* sh_send = nullptr;  // [n_out][T][B]
  float2* sh_recv = nullptr;  // [sh_nloc][T][B]
  uint32_t* sh_view = nullptr;  // device: first[n_out] | count[n_out] | xjob[n_out], first[o] = o - sh_o0
  // MIMO on the tensor cores (mimo_tc.cuh): bin-major operand copies and a one-slot-per-output plan view
  bool tc_on = false;          // buffers exist (MIMO mode, max_blocks >= tc_min_blocks, not disabled)
  bool tc_dirty = true;        // Hpack must be rebuilt from the current filter matrix
  uint32_t tc_min_blocks = 16; // calls with fewer blocks use the streaming SIMT MAC
  uint32_t tc_P2 = 1, tc_P2log = 0, tc_G = 0, tc_nog = 0;
  uint64_t tc_xbin_max = 0;  // float2 elements per bin of tc_xb at N = 64
  float4* tc_hpack = nullptr;
  float2* tc_xb = nullptr;
  const float2** tc_ftab_h = nullptr;  // pinned staging [n_out][n_in]
  const float2** tc_ftab_d = nullptr;
  uint32_t* tc_fparts_h = nullptr;
  uint32_t* tc_fparts_d = nullptr;
  uint32_t* tc_view = nullptr;  // device: first[n_out] | count[n_out] | xjob[n_out]
  int* tc_status = nullptr;    // device view of tc_status_h
  int* tc_status_h = nullptr;  // mapped pinned word the kernel sets when a barrier wait times out
  unsigned long long* tc_trace = nullptr;  // optional per-CTA role cycle counters (bbx_engine_tensor_trace)
  uint32_t tc_trace_ctas = 0;
  uint64_t tc_launches = 0;
};

namespace {

// persistent grid of the radix-8 kernels: enough CTAs to fill the machine, never more than there are items
template <typename K>
uint32_t persistent_grid(K kernel, int threads, size_t smem, uint32_t nitems) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return std::max(1u, std::min(nitems, (uint32_t)(kNumSMs * per_sm)));
}

template <int M>
void launch_rfft_t(const float* src, uint64_t ch_stride, uint32_t win_stride, float2* dst, uint64_t dst_ch_stride, uint32_t R,
                   uint32_t slot0, const float2* tw, float scale, uint32_t nch, uint32_t T, cudaStream_t st) {
  constexpr int FPB = FftCfg<M>::FPB;
  if constexpr (FftCfg<M>::R == 8) {
    static uint32_t per_sm_grid = 0;  // occupancy query once per size
    if (!per_sm_grid) per_sm_grid = persistent_grid(k_rfft8<M>, FftCfg<M>::NT * FPB, 0, 1u << 30);
    const uint32_t nitems = ceil_div(nch, FPB) * T;
    k_rfft8<M><<<std::min(nitems, per_sm_grid), dim3(FftCfg<M>::NT, FPB), 0, st>>>(src, ch_stride, win_stride, dst, dst_ch_stride, R,
                                                                                  slot0, tw, scale, nch, T);
    return;
  }
  k_rfft<M><<<dim3(ceil_div(nch, FPB), T), dim3(FftCfg<M>::NT, FPB), 0, st>>>(src, ch_stride, win_stride, dst, dst_ch_stride, R,
                                                                              slot0, tw, scale, nch);
}

int launch_rfft(uint32_t B, const float* src, uint64_t ch_stride, uint32_t win_stride, float2* dst, uint64_t dst_ch_stride,
                uint32_t R, uint32_t slot0, const float2* tw, float scale, uint32_t nch, uint32_t T, cudaStream_t st) {
  switch (B) {
    case 64: launch_rfft_t<64>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 128: launch_rfft_t<128>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 256: launch_rfft_t<256>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 512: launch_rfft_t<512>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 1024: launch_rfft_t<1024>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 2048: launch_rfft_t<2048>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 4096: launch_rfft_t<4096>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    default: set_error("unsupported block size %u", B); return BBX_ERR_UNSUPPORTED;
  }
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

// view of the tensor-core MIMO result: output o = slot o, one slot per job, no crossfade
PlanView tc_plan_view(const bbx_engine* e) {
  PlanView v;
  v.job_slot_first = e->tc_view;
  v.job_slot_count = e->tc_view + e->n_out;
  v.xjob = e->tc_view + 2 * (size_t)e->n_out;
  return v;
}

PlanView peer_plan_view(const bbx_engine* e) {
  PlanView v;
  v.job_slot_first = e->px_view;
  v.job_slot_count = e->px_view + e->n_out;
  v.xjob = e->px_view + 2 * (size_t)e->n_out;
  return v;
}

PlanView shard_plan_view(const bbx_engine* e) {
  PlanView v;
  v.job_slot_first = e->sh_view;
  v.job_slot_count = e->sh_view + e->n_out;
  v.xjob = e->sh_view + 2 * (size_t)e->n_out;
  return v;
}

template <int M>
void launch_irfft_t(bbx_engine* e, uint32_t T, uint32_t n_first, bool tc, cudaStream_t st) {
  const float2* ypart = e->ypart;
  const float* nyq_part = e->nyq_part;
  const uint32_t wpos = e->wpos;
  constexpr int FPB = FftCfg<M>::FPB;
  constexpr size_t smem = sizeof(float2) * (size_t)FPB * (M + FftCfg<M>::MP);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_irfft<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e->sh_world > 1 || e->comm) {
    // reduced spectra of the local outputs, [local output][t][M]; peer mode: [local output][source rank][t][M], the
    // kernel adds the `world` slots of an output in rank order
    const PlanView v = e->px_on ? peer_plan_view(e) : shard_plan_view(e);
    const float2* spectra = e->px_on ? (const float2*)e->px_mem + (uint64_t)(e->px_epoch & 1u) * e->px_half : e->sh_recv;
    if constexpr (FftCfg<M>::R == 8) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(k_irfft8<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const uint32_t nitems = ceil_div(e->sh_nloc, FPB) * T;
      k_irfft8<M><<<std::min(nitems, (uint32_t)kNumSMs * 2), dim3(FftCfg<M>::NT, FPB), smem, st>>>(
          spectra, e->max_slots, v, v, 0, e->tw, e->ybuf, e->Rd, e->wpos, e->sh_nloc, nullptr, (uint64_t)M, (uint64_t)T * M,
          e->sh_o0, T);
      return;
    }
    k_irfft<M><<<dim3(ceil_div(e->sh_nloc, FPB), T), dim3(FftCfg<M>::NT, FPB), smem, st>>>(
        spectra, e->max_slots, v, v, 0, e->tw, e->ybuf, e->Rd, e->wpos, e->sh_nloc, nullptr, (uint64_t)M, (uint64_t)T * M,
        e->sh_o0);
    return;
  }
  const PlanView first = tc ? tc_plan_view(e) : e->plan_first.view();
  const PlanView steady = tc ? tc_plan_view(e) : e->plan_steady.view();
  if constexpr (FftCfg<M>::R == 8) {
    static uint32_t per_sm_grid = 0;
    if (!per_sm_grid) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(k_irfft8<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      per_sm_grid = persistent_grid(k_irfft8<M>, FftCfg<M>::NT * FPB, smem, 1u << 30);
    }
    const uint32_t nitems = ceil_div(e->n_streams, FPB) * T;
    k_irfft8<M><<<std::min(nitems, per_sm_grid), dim3(FftCfg<M>::NT, FPB), smem, st>>>(
        ypart, e->max_slots, first, steady, n_first, e->tw, e->ybuf, e->Rd, wpos, e->n_streams, tc ? nullptr : nyq_part,
        (uint64_t)e->max_slots * M, (uint64_t)M, 0u, T);
    return;
  }
  k_irfft<M><<<dim3(ceil_div(e->n_streams, FPB), T), dim3(FftCfg<M>::NT, FPB), smem, st>>>(
      ypart, e->max_slots, first, steady, n_first, e->tw, e->ybuf, e->Rd, wpos, e->n_streams,
      tc ? nullptr : nyq_part, (uint64_t)e->max_slots * M, (uint64_t)M, 0u);
}

int launch_irfft(bbx_engine* e, uint32_t T, uint32_t n_first, bool tc, cudaStream_t st) {
  switch (e->B) {
    case 64: launch_irfft_t<64>(e, T, n_first, tc, st); break;
    case 128: launch_irfft_t<128>(e, T, n_first, tc, st); break;
    case 256: launch_irfft_t<256>(e, T, n_first, tc, st); break;
    case 512: launch_irfft_t<512>(e, T, n_first, tc, st); break;
    case 1024: launch_irfft_t<1024>(e, T, n_first, tc, st); break;
    case 2048: launch_irfft_t<2048>(e, T, n_first, tc, st); break;
    case 4096: launch_irfft_t<4096>(e, T, n_first, tc, st); break;
    default: set_error("unsupported block size %u", e->B); return BBX_ERR_UNSUPPORTED;
  }
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

template <int THREADS>
void launch_mac_t(bbx_engine* e, const MacPlan& pl, uint32_t t0, uint32_t nt, cudaStream_t st) {
  const uint32_t halfB = e->B / 2;
  dim3 grid(pl.n_ctas, halfB / THREADS, nt);
  float4* yp = (float4*)e->ypart + (uint64_t)t0 * e->max_slots * halfB;
  float* nq = e->nyq_part + (uint64_t)t0 * e->max_slots;
  const MacSeg* segs = pl.segs();
  const uint32_t* cta = pl.cta_seg_begin();
  const float4* fdl = (const float4*)e->fdl;
  const float keep = e->mac_l2_keep;
  // FDL rows are read once per step only when every input feeds one path; shared inputs (ROUTED fan-out, MIMO)
  // re-read them within the step and must keep normal L2 priority
  const int px = (e->mode == BBX_MODE_PER_CHANNEL) ? 1 : 0;
#define BBX_MAC_LAUNCH(U, OCC)                                                                                            \
  do {                                                                                                                    \
    if (keep > 0.f)                                                                                                       \
      k_fdl_mac<U, THREADS, OCC, true><<<grid, THREADS, 0, st>>>(segs, cta, fdl, yp, nq, halfB, e->R, e->head, t0, e->max_slots, keep, px); \
    else                                                                                                                  \
      k_fdl_mac<U, THREADS, OCC, false><<<grid, THREADS, 0, st>>>(segs, cta, fdl, yp, nq, halfB, e->R, e->head, t0, e->max_slots, keep, px); \
  } while (0)
  // resident CTAs per SM <-> loads in flight per thread: fewer, fatter CTAs unroll deeper
  switch (pl.occ) {
    case 1: BBX_MAC_LAUNCH(16, 1); break;
    case 2: BBX_MAC_LAUNCH(8, 2); break;
    case 3: BBX_MAC_LAUNCH(6, 3); break;
    default: BBX_MAC_LAUNCH(4, 4); break;
  }
#undef BBX_MAC_LAUNCH
}

template <int TT, int THREADS>
void launch_mac_tb_t(bbx_engine* e, const MacPlan& pl, uint32_t t0, uint32_t nt, cudaStream_t st) {
  const uint32_t ncol = e->B / THREADS, ntile = ceil_div(nt, TT);
  dim3 grid(ncol * ntile, pl.n_ctas);
  float2* yp = e->ypart + (uint64_t)t0 * e->max_slots * e->B;
  k_fdl_mac_tb<TT, THREADS, 8><<<grid, THREADS, 0, st>>>(pl.segs(), pl.cta_seg_begin(), e->fdl, yp, e->B, e->R, e->head, t0, nt,
                                                        ncol, e->max_slots);

}

template <int TT>
void launch_mac_tb(bbx_engine* e, const MacPlan& pl, uint32_t t0, uint32_t nt, cudaStream_t st) {
  if (e->B >= 256) launch_mac_tb_t<TT, 256>(e, pl, t0, nt, st);
  else if (e->B == 128) launch_mac_tb_t<TT, 128>(e, pl, t0, nt, st);
  else launch_mac_tb_t<TT, 64>(e, pl, t0, nt, st);
}

// the time-batched kernel pays a window fill of TT-1 rows per term: only worth it for long filters and enough block-steps
bool mac_uses_time_batching(const bbx_engine* e, const MacPlan& pl, uint32_t nt) {
  const uint32_t tb = e->mac_time_tile;  // 0: streaming only
  return tb && nt >= tb / 2 && pl.n_terms && pl.total_rows / pl.n_terms >= 2 * tb;
}

int launch_mac(bbx_engine* e, const MacPlan& pl, uint32_t t0, uint32_t nt) {
  if (pl.n_ctas == 0 || nt == 0) return BBX_OK;
  cudaStream_t st = e->stream;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e->profile_mac) {
    if (e->mac_events_used + 2 > e->mac_events.size()) {
      size_t old = e->mac_events.size();
      e->mac_events.resize(old + 64);
      for (size_t i = old; i < e->mac_events.size(); i++) BBX_CUDA_TRY(cudaEventCreate(&e->mac_events[i]));
    }
    ev0 = e->mac_events[e->mac_events_used++];
    ev1 = e->mac_events[e->mac_events_used++];
  }
  const uint32_t halfB = e->B / 2;
  // the time-batched kernel pays a window fill of TT-1 rows per term: only worth it for long filters
  const uint32_t tb = e->mac_time_tile;  // 0: streaming only
  const bool use_tb = mac_uses_time_batching(e, pl, nt);
  if (use_tb) {
    // Nyquist sums of column 0 (the streaming kernel accumulates them inline): a few hundred latency-bound warps,
    // forked onto the side stream so they run underneath the MAC instead of after it
    BBX_CUDA_TRY(cudaEventRecord(e->ev_fork, st));
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_aux, e->ev_fork, 0));
    k_nyq_mac<<<dim3(pl.n_ctas, ceil_div(nt, 32)), 32, 0, e->s_aux>>>(pl.segs(), pl.cta_seg_begin(), e->fdl,
                                                                     e->nyq_part + (uint64_t)t0 * e->max_slots, e->B, e->R,
                                                                     e->head, t0, nt, e->max_slots);
    BBX_CUDA_TRY(cudaGetLastError());
    BBX_CUDA_TRY(cudaEventRecord(e->ev_join, e->s_aux));
    e->launches++;
  }
  if (ev0) BBX_CUDA_TRY(cudaEventRecord(ev0, st));
  if (use_tb) {
    if (tb == 32) launch_mac_tb<32>(e, pl, t0, nt, st);
    else launch_mac_tb<16>(e, pl, t0, nt, st);
  } else if (halfB >= 256) launch_mac_t<256>(e, pl, t0, nt, st);
  else if (halfB == 128) launch_mac_t<128>(e, pl, t0, nt, st);
  else if (halfB == 64) launch_mac_t<64>(e, pl, t0, nt, st);
  else launch_mac_t<32>(e, pl, t0, nt, st);
  BBX_CUDA_TRY(cudaGetLastError());
  if (ev1) BBX_CUDA_TRY(cudaEventRecord(ev1, st));
  e->launches++;
  if (use_tb) BBX_CUDA_TRY(cudaStreamWaitEvent(st, e->ev_join, 0));
  e->mac_launches++;
  e->mac_units += (uint64_t)e->n_streams * nt;
  // 16 P K + 16 K + (bytes_in + bytes_out) B per channel-block, K = B + 1
  const uint64_t K = e->B + 1;
  e->mac_bytes += (uint64_t)nt * (16ull * pl.total_rows * K + 16ull * K * e->n_streams +
                                  (uint64_t)e->B * (fmt_bytes(e->last_infmt) * e->n_in + fmt_bytes(e->last_outfmt) * e->n_out));
  return BBX_OK;
}

// The pinned plan/route staging is rewritten by the host; wait until the previous upload has read it.
int wait_uploads(bbx_engine* e) {
  if (e->upload_pending) {
    BBX_CUDA_TRY(cudaEventSynchronize(e->ev_upload));
    e->upload_pending = false;
  }
  return BBX_OK;
}
int mark_upload(bbx_engine* e) {
  BBX_CUDA_TRY(cudaEventRecord(e->ev_upload, e->stream));
  e->upload_pending = true;
  return BBX_OK;
}

int staging_alloc(Staging& s, size_t bytes) {
  s.bytes = bytes;
  for (int i = 0; i < Staging::kDepth; i++) {
    BBX_CUDA_TRY(cudaHostAlloc((void**)&s.h[i], bytes, cudaHostAllocDefault));
    memset(s.h[i], 0, bytes);
    BBX_CUDA_TRY(cudaEventCreateWithFlags(&s.ev[i], cudaEventDisableTiming));
  }
  return BBX_OK;
}
void staging_free(Staging& s) {
  for (int i = 0; i < Staging::kDepth; i++) {
    if (s.h[i]) cudaFreeHost(s.h[i]);
    if (s.ev[i]) cudaEventDestroy(s.ev[i]);
    s.h[i] = nullptr;
    s.ev[i] = nullptr;
  }
}
// next slot to fill; blocks only while the copy that last read this slot is still queued
int staging_acquire(Staging& s, uint8_t** out) {
  if (s.pending[s.cur]) {
    BBX_CUDA_TRY(cudaEventSynchronize(s.ev[s.cur]));
    s.pending[s.cur] = false;
  }
  *out = s.h[s.cur];
  return BBX_OK;
}
int staging_commit(Staging& s, void* dst, cudaStream_t st) {
  BBX_CUDA_TRY(cudaMemcpyAsync(dst, s.h[s.cur], s.bytes, cudaMemcpyHostToDevice, st));
  BBX_CUDA_TRY(cudaEventRecord(s.ev[s.cur], st));
  s.pending[s.cur] = true;
  s.cur = (s.cur + 1) % Staging::kDepth;
  return BBX_OK;
}

// Build a MAC plan from a job list into a pinned staging slot and enqueue its upload.
int build_plan(bbx_engine* e, MacPlan& pl, const std::vector<std::vector<JobTerm>>& jobs, const std::vector<uint32_t>& xjob) {
  uint32_t total = 0, nterms = 0;
  for (auto& j : jobs)
    for (auto& tm : j)
      if (tm.f) {
        total += tm.f->P;
        nterms++;
      }
  pl.n_terms = nterms;
  {
    int wrc = staging_acquire(pl.stg, &pl.h_blob);
    if (wrc) return wrc;
  }
  MacSeg* segs = (MacSeg*)(pl.h_blob + pl.off_segs);
  uint32_t* cta = (uint32_t*)(pl.h_blob + pl.off_cta);
  uint32_t* jfirst = (uint32_t*)(pl.h_blob + pl.off_first);
  uint32_t* jcount = (uint32_t*)(pl.h_blob + pl.off_count);
  uint32_t* xj = (uint32_t*)(pl.h_blob + pl.off_xjob);
  BBX_REQUIRE(jobs.size() <= e->max_jobs, "internal: too many jobs");
  pl.n_jobs = (uint32_t)jobs.size();
  pl.total_rows = total;
  for (uint32_t s = 0; s < e->n_streams; s++) xj[s] = s < xjob.size() ? xjob[s] : kNoJob;
  if (total == 0) {
    for (uint32_t j = 0; j < jobs.size(); j++) jfirst[j] = jcount[j] = 0;
    pl.n_ctas = 0;
    pl.n_slots = 0;
  } else {
    // even split of the flattened row space; small problems get fewer, fatter CTAs
    const uint32_t min_rows = 4;
    // short filters (<= 8 partitions per term, e.g. the 64x64 MIMO matrix) cannot fill a 16-deep unroll: cut the
    // plan for two resident CTAs per SM with 8 rows in flight each instead
    pl.occ = (nterms && total / nterms <= 8 && e->mac_occ < 2) ? 2u : e->mac_occ;
    uint32_t G = std::min(kNumSMs * pl.occ, std::max(1u, total / min_rows));
    uint32_t rpc = ceil_div(total, G);
    G = ceil_div(total, rpc);
    uint32_t nseg = 0, nslot = 0, row = 0, cur_cta = 0;
    cta[0] = 0;
    int run_job = -1;  // job of the open run inside the current CTA
    for (uint32_t j = 0; j < jobs.size(); j++) {
      jfirst[j] = nslot;
      uint32_t before = nslot;
      bool job_has_run = false;
      for (auto& tm : jobs[j]) {
        if (!tm.f) continue;
        uint32_t p = 0;
        while (p < tm.f->P) {
          uint32_t cta_end = (cur_cta + 1) * rpc;
          if (row == cta_end) {  // move to the next CTA
            cur_cta++;
            cta[cur_cta] = nseg;
            run_job = -1;
            continue;
          }
          uint32_t np = std::min(tm.f->P - p, cta_end - row);
          BBX_REQUIRE(nseg < e->max_segs, "internal: MAC plan overflow (segments)");
          MacSeg& sg = segs[nseg];
          sg.H = (const float4*)tm.f->H;
          sg.fdl_ch = tm.input;
          sg.p0 = p;
          sg.np = np;
          sg.pad = 0;
          if (run_job == (int)j) {
            // continue the open run: the previous segment no longer writes
            segs[nseg - 1].flags &= ~2u;
            sg.flags = 2u;
            sg.slot = segs[nseg - 1].slot;
          } else {
            BBX_REQUIRE(nslot < e->max_slots, "internal: MAC plan overflow (slots)");
            sg.flags = 1u | 2u;
            sg.slot = nslot++;
            run_job = (int)j;
            job_has_run = true;
          }
          nseg++;
          p += np;
          row += np;
        }
      }
      (void)job_has_run;
      jcount[j] = nslot - before;
    }
    for (uint32_t c = cur_cta + 1; c <= G; c++) cta[c] = nseg;
    pl.n_ctas = G;
    pl.n_slots = nslot;
  }
  pl.valid = true;
  return staging_commit(pl.stg, pl.d_blob, e->stream);
}

int alloc_plan(bbx_engine* e, MacPlan& pl) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  pl.off_segs = take(sizeof(MacSeg) * e->max_segs);
  pl.off_cta = take(sizeof(uint32_t) * (e->max_ctas + 2));
  pl.off_first = take(sizeof(uint32_t) * e->max_jobs);
  pl.off_count = take(sizeof(uint32_t) * e->max_jobs);
  pl.off_xjob = take(sizeof(uint32_t) * std::max(1u, e->n_streams));
  pl.blob_bytes = off;
  {
    int src = staging_alloc(pl.stg, off);
    if (src) return src;
  }
  BBX_CUDA_TRY(cudaMalloc((void**)&pl.d_blob, off));
  BBX_CUDA_TRY(cudaMemset(pl.d_blob, 0, off));
  return BBX_OK;
}

// jobs for the given choice of filter per path; MIMO groups the paths of one output into one job
void make_jobs(const bbx_engine* e, bool use_pending, std::vector<std::vector<JobTerm>>& jobs) {
  jobs.clear();
  auto pick = [&](const PathState& p) { return (use_pending && p.has_pending) ? p.pend : p.cur; };
  if (e->mode == BBX_MODE_MIMO) {
    jobs.resize(e->n_out);
    for (uint32_t o = 0; o < e->n_out; o++)
      for (uint32_t i = 0; i < e->n_in; i++) {
        const PathState& p = e->paths[(size_t)o * e->n_in + i];
        const bbx_filter* f = pick(p);
        if (f) jobs[o].push_back({f, i});
      }
  } else {
    jobs.resize(e->n_paths);
    for (uint32_t k = 0; k < e->n_paths; k++) {
      const PathState& p = e->paths[k];
      const bbx_filter* f = pick(p);
      if (f) jobs[k].push_back({f, p.input});
    }
  }
}

int upload_routes(bbx_engine* e, bool first_block_transition) {
  {
    int wrc = staging_acquire(e->route_stg, &e->h_route);
    if (wrc) return wrc;
  }
  uint32_t* ofirst = (uint32_t*)(e->h_route + e->roff_first);
  uint32_t* rstream = (uint32_t*)(e->h_route + e->roff_stream);
  float* gain = (float*)(e->h_route + e->roff_gain);
  double* dcur = (double*)(e->h_route + e->roff_dcur);
  double* dold = (double*)(e->h_route + e->roff_dold);
  uint32_t* flags = (uint32_t*)(e->h_route + e->roff_flags);
  uint32_t* icur = (uint32_t*)(e->h_route + e->roff_icur);
  uint32_t* iold = (uint32_t*)(e->h_route + e->roff_iold);
  if (e->mode == BBX_MODE_MIMO) {
    for (uint32_t o = 0; o < e->n_out; o++) {
      ofirst[o] = o;
      rstream[o] = o;
      gain[o] = 1.0f;
      dcur[o] = dold[o] = 0.0;
      icur[o] = iold[o] = 0;
      flags[o] = 0;
    }
    ofirst[e->n_out] = e->n_out;
    // sharded: PCM channel c of this rank is output sh_o0 + c
    for (uint32_t c = 0; c < e->n_out_pcm && e->n_out_pcm < e->n_out; c++) rstream[c] = e->sh_o0 + c;
  } else {
    uint32_t n = 0;
    for (uint32_t o = 0; o < e->n_out; o++) {
      ofirst[o] = n;
      for (uint32_t k = 0; k < e->n_paths; k++)
        if (e->paths[k].output == o) rstream[n++] = k;
    }
    ofirst[e->n_out] = n;
    for (uint32_t k = 0; k < e->n_paths; k++) {
      const PathState& p = e->paths[k];
      gain[k] = p.gain;
      bool sw = first_block_transition && p.has_pending;
      dcur[k] = sw ? p.pend_delay : p.delay;
      dold[k] = p.delay;
      icur[k] = (uint32_t)dcur[k] % e->Rd;
      iold[k] = (uint32_t)dold[k] % e->Rd;
      flags[k] = (sw && p.xfade && p.pend_delay != p.delay) ? 1u : 0u;
    }
  }
  // per-route entries in mixdown order (what k_pcm_out reads)
  {
    RouteEntry* en = (RouteEntry*)(e->h_route + e->roff_entry);
    const uint32_t nroutes = ofirst[e->n_out_pcm < e->n_out ? e->n_out_pcm : e->n_out];
    e->n_routes_pcm = nroutes;
    for (uint32_t r = 0; r < nroutes; r++) {
      const uint32_t st = rstream[r];
      en[r].stream = st;
      en[r].gain = gain[st];
      en[r].icur = icur[st];
      en[r].iold = iold[st];
      en[r].flags = flags[st];
      en[r].pad = 0;
      en[r].dcur = dcur[st];
      en[r].dold = dold[st];
    }
  }
  return staging_commit(e->route_stg, e->d_route, e->stream);
}

RouteView route_view(const bbx_engine* e) {
  RouteView v;
  v.out_first = (const uint32_t*)(e->d_route + e->roff_first);
  v.entry = (const RouteEntry*)(e->d_route + e->roff_entry);
  return v;
}

bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

// ---- MIMO on the tensor cores ----
int tc_alloc(bbx_engine* e) {
  uint32_t P2 = 1, lg = 0;
  while (P2 < e->Pmax) {
    P2 <<= 1;
    lg++;
  }
  e->tc_P2 = P2;
  e->tc_P2log = lg;
  const uint32_t Kc = ceil_div(e->n_in * P2, (uint32_t)kTcChunk) * kTcChunk;  // complex K, padded to whole chunks
  e->tc_G = Kc / 2;
  e->tc_nog = ceil_div(e->n_out, 64u);
  // pre-staged FDL runs: per bin [column tile][chunk][hi | lo][rows of the chunk][seg], sized for N = 64
  {
    const uint32_t ninp = lg < 4 ? (16u >> lg) : 1u, seg = lg < 4 ? ((kTcNmax + P2) & ~1u) : kTcNmax + 16;
    e->tc_xbin_max = (uint64_t)ceil_div(e->Tmax, (uint32_t)kTcNmax) * (Kc / kTcChunk) * 2 * ninp * seg;
  }
  const size_t hbytes = sizeof(float4) * (size_t)e->tc_nog * e->B * e->tc_G * 64;
  const size_t xbytes = sizeof(float2) * (size_t)e->B * e->tc_xbin_max;
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_hpack, hbytes));
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_xb, xbytes));
  const size_t npaths = (size_t)e->n_out * e->n_in;
  BBX_CUDA_TRY(cudaHostAlloc((void**)&e->tc_ftab_h, sizeof(float2*) * npaths, cudaHostAllocDefault));
  BBX_CUDA_TRY(cudaHostAlloc((void**)&e->tc_fparts_h, sizeof(uint32_t) * npaths, cudaHostAllocDefault));
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_ftab_d, sizeof(float2*) * npaths));
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_fparts_d, sizeof(uint32_t) * npaths));
  std::vector<uint32_t> view(3 * (size_t)e->n_out);
  for (uint32_t o = 0; o < e->n_out; o++) {
    view[o] = o;
    view[e->n_out + o] = 1;
    view[2 * (size_t)e->n_out + o] = kNoJob;
  }
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_view, sizeof(uint32_t) * view.size()));
  BBX_CUDA_TRY(cudaMemcpy(e->tc_view, view.data(), sizeof(uint32_t) * view.size(), cudaMemcpyHostToDevice));
  BBX_CUDA_TRY(cudaHostAlloc((void**)&e->tc_status_h, sizeof(int), cudaHostAllocMapped));
  *e->tc_status_h = 0;
  BBX_CUDA_TRY(cudaHostGetDevicePointer((void**)&e->tc_status, e->tc_status_h, 0));
  BBX_CUDA_TRY(cudaFuncSetAttribute(k_mimo_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemMax));
  BBX_CUDA_TRY(cudaFuncSetAttribute(k_mimo_tc<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemMax));
  BBX_CUDA_TRY(cudaFuncSetAttribute(k_mimo_tc<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemMax));
  e->tc_on = true;
  e->tc_dirty = true;
  return BBX_OK;
}

// rebuild the bin-major filter operand from the current filter matrix
int tc_pack_filters(bbx_engine* e) {
  int rc = wait_uploads(e);
  if (rc) return rc;
  const size_t npaths = (size_t)e->n_out * e->n_in;
  for (size_t k = 0; k < npaths; k++) {
    const bbx_filter* f = e->paths[k].cur;
    e->tc_ftab_h[k] = f ? f->H : nullptr;
    e->tc_fparts_h[k] = f ? f->P : 0;
  }
  BBX_CUDA_TRY(cudaMemcpyAsync(e->tc_ftab_d, e->tc_ftab_h, sizeof(float2*) * npaths, cudaMemcpyHostToDevice, e->stream));
  BBX_CUDA_TRY(cudaMemcpyAsync(e->tc_fparts_d, e->tc_fparts_h, sizeof(uint32_t) * npaths, cudaMemcpyHostToDevice, e->stream));
  if ((rc = mark_upload(e))) return rc;
  k_mimo_pack_h<<<dim3(e->B / 32, e->tc_G, e->tc_nog), 256, 0, e->stream>>>(e->tc_ftab_d, e->tc_fparts_d, e->tc_hpack, e->B, e->n_in,
                                                                        e->n_out, e->tc_P2log, e->tc_G);
  BBX_CUDA_TRY(cudaGetLastError());
  e->launches++;
  e->tc_dirty = false;
  return BBX_OK;
}

// the T block-steps of one call: FDL rows -> bin-major X, then one GEMM per bin
int launch_mimo_tc(bbx_engine* e, uint32_t T) {
  cudaStream_t st = e->stream;
  uint32_t N = 16, Nlog = 4;
  while (N < T && N < (uint32_t)kTcNmax) {
    N <<= 1;
    Nlog++;
  }
  const uint32_t ntiles = ceil_div(T, N);
  const uint32_t nchunk = e->tc_G / (kTcChunk / 2);
  const uint32_t ninp = e->tc_P2log < 4 ? (16u >> e->tc_P2log) : 1u;
  const uint32_t seg = e->tc_P2log < 4 ? ((N + e->tc_P2) & ~1u) : N + 16;
  const uint64_t xbin = (uint64_t)ntiles * nchunk * 2 * ninp * seg;
  k_mimo_pack_x<<<dim3(e->B / 32, ntiles * nchunk * ninp, ceil_div(seg, 32u)), 256, 0, st>>>(e->fdl, e->tc_xb, e->B, e->R, e->head,
                                                                                         e->n_in, e->tc_P2log, T, N, nchunk, xbin);
  BBX_CUDA_TRY(cudaGetLastError());
  e->launches++;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e->profile_mac) {
    if (e->mac_events_used + 2 > e->mac_events.size()) {
      size_t old = e->mac_events.size();
      e->mac_events.resize(old + 64);
      for (size_t i = old; i < e->mac_events.size(); i++) BBX_CUDA_TRY(cudaEventCreate(&e->mac_events[i]));
    }
    ev0 = e->mac_events[e->mac_events_used++];
    ev1 = e->mac_events[e->mac_events_used++];
  }
  MimoTcArgs a;
  a.hpack = e->tc_hpack;
  a.xb = e->tc_xb;
  a.ypart = e->ypart;
  a.status = e->tc_status;
  a.B = e->B;
  a.n_in = e->n_in;
  a.n_out = e->n_out;
  a.P2log = e->tc_P2log;
  a.G = e->tc_G;
  a.T = T;
  a.slot_stride = e->max_slots;
  a.xbin = xbin;
  a.trace = nullptr;
  if (ev0) BBX_CUDA_TRY(cudaEventRecord(ev0, st));
  // raw ring: 8 KB of H + hi and lo FDL runs per chunk (see mimo_tc.cuh); as many stages as fit, an even number
  {
    a.raw_stage_bytes = (8192 + 2 * ninp * seg * 8 + 127) & ~127u;
    uint32_t nst = (kTcSmemMax - kTcOffRaw) / a.raw_stage_bytes;
    nst = std::min(nst, (uint32_t)kTcRawStagesMax) & ~1u;
    a.raw_stages = nst;
  }
  const uint32_t smem = kTcOffRaw + a.raw_stages * a.raw_stage_bytes;
  const dim3 grid(e->B / kTcBins, e->tc_nog, ntiles);
  if (e->tc_trace && grid.x * grid.y * grid.z <= e->tc_trace_ctas) a.trace = e->tc_trace;
  if (Nlog == 4) k_mimo_tc<4><<<grid, kTcThreads, smem, st>>>(a);
  else if (Nlog == 5) k_mimo_tc<5><<<grid, kTcThreads, smem, st>>>(a);
  else k_mimo_tc<6><<<grid, kTcThreads, smem, st>>>(a);
  BBX_CUDA_TRY(cudaGetLastError());
  if (ev1) BBX_CUDA_TRY(cudaEventRecord(ev1, st));
  e->launches++;
  e->tc_launches++;
  e->mac_launches++;
  e->mac_units += (uint64_t)e->n_streams * T;
  const uint64_t K = e->B + 1;
  e->mac_bytes += (uint64_t)T * (16ull * e->plan_steady.total_rows * K + 16ull * K * e->n_streams +
                                 (uint64_t)e->B * (fmt_bytes(e->last_infmt) * e->n_in + fmt_bytes(e->last_outfmt) * e->n_out));
  return BBX_OK;
}

// 0 always means "leave as is" (library default at creation)
void apply_tuning(bbx_engine* e, uint32_t ctas_per_sm, uint32_t l2_keep_16ths, uint32_t time_tile) {
  if (ctas_per_sm) e->mac_occ = std::min(ctas_per_sm, 4u);
  if (l2_keep_16ths) e->mac_l2_keep = (l2_keep_16ths > 16u) ? 0.f : l2_keep_16ths / 16.0f;  // > 16: hints off
  if (time_tile) e->mac_time_tile = (time_tile == 16 || time_tile == 32) ? time_tile : 0;     // 1: streaming only
  e->steady_dirty = true;
}

}  // namespace

extern "C" {

int bbx_engine_create(const bbx_config* cfg, bbx_engine** out) {
  BBX_REQUIRE(cfg && out, "bbx_engine_create: null argument");
  int rc = require_device();
  if (rc) return rc;
  BBX_REQUIRE(is_pow2(cfg->block_size) && cfg->block_size >= 64 && cfg->block_size <= 4096,
              "block_size %u must be a power of two in [64, 4096]", cfg->block_size);
  BBX_REQUIRE(cfg->n_inputs > 0, "n_inputs must be > 0");
  BBX_REQUIRE(cfg->mode >= BBX_MODE_PER_CHANNEL && cfg->mode <= BBX_MODE_MIMO, "bad mode %d", cfg->mode);
  bbx_engine* e = new bbx_engine();
  e->cfg = *cfg;
  e->device = cfg->device;
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  e->B = cfg->block_size;
  e->Pmax = std::max(1u, cfg->max_partitions);
  e->n_in = cfg->n_inputs;
  e->mode = cfg->mode;
  if (e->mode == BBX_MODE_PER_CHANNEL) {
    e->n_out = e->n_in;
    e->n_paths = e->n_in;
    e->n_streams = e->n_in;
  } else if (e->mode == BBX_MODE_ROUTED) {
    BBX_REQUIRE(cfg->n_outputs > 0 && cfg->n_paths > 0, "ROUTED mode needs n_outputs and n_paths");
    e->n_out = cfg->n_outputs;
    e->n_paths = cfg->n_paths;
    e->n_streams = cfg->n_paths;
  } else {
    BBX_REQUIRE(cfg->n_outputs > 0, "MIMO mode needs n_outputs");
    BBX_REQUIRE(cfg->max_delay == 0, "MIMO mode has no per-path delay (frequency-domain mixdown)");
    e->n_out = cfg->n_outputs;
    e->n_paths = e->n_in * e->n_out;
    e->n_streams = e->n_out;
    if (cfg->mimo_shard_world > 1) {
      BBX_REQUIRE(cfg->mimo_shard_rank < cfg->mimo_shard_world, "mimo_shard_rank %u out of range", cfg->mimo_shard_rank);
      BBX_REQUIRE(e->n_out % cfg->mimo_shard_world == 0, "n_outputs %u is not a multiple of mimo_shard_world %u", e->n_out,
                  cfg->mimo_shard_world);
      e->sh_world = cfg->mimo_shard_world;
      e->sh_rank = cfg->mimo_shard_rank;
    }
  }
  BBX_REQUIRE(e->mode == BBX_MODE_MIMO || cfg->mimo_shard_world <= 1, "mimo_shard_world needs MIMO mode");
  e->sh_nloc = e->n_out / e->sh_world;
  e->sh_o0 = e->sh_rank * e->sh_nloc;
  e->n_out_pcm = (e->sh_world > 1) ? e->sh_nloc : e->n_out;
  e->Tmax = std::max(1u, cfg->max_blocks);
  e->R = e->Pmax + e->Tmax - 1;
  uint32_t need = cfg->max_delay + 14 + (e->Tmax + 1) * e->B;
  e->Rd = cfg->ring_length ? cfg->ring_length : ceil_div(need, e->B) * e->B;
  BBX_REQUIRE(e->Rd >= cfg->max_delay + 14 + e->Tmax * e->B, "ring_length %u too short (need >= %u)", e->Rd,
              cfg->max_delay + 14 + e->Tmax * e->B);
  e->xstride = (e->Tmax + 1) * e->B;
  e->paths.resize(e->n_paths);
  for (uint32_t k = 0; k < e->n_paths; k++) {
    PathState& p = e->paths[k];
    if (e->mode == BBX_MODE_MIMO) {
      p.output = k / e->n_in;
      p.input = k % e->n_in;
    } else if (e->mode == BBX_MODE_PER_CHANNEL) {
      p.input = p.output = k;
    }
  }
  BBX_CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  BBX_CUDA_TRY(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
  BBX_CUDA_TRY(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
  BBX_CUDA_TRY(cudaStreamCreateWithFlags(&e->s_aux, cudaStreamNonBlocking));
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) {
    BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming));
    BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_comp[i], cudaEventDisableTiming));
    BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_d2h[i], cudaEventDisableTiming));
  }
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join_in, cudaEventDisableTiming));
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join_out, cudaEventDisableTiming));
  BBX_CUDA_TRY(cudaEventCreate(&e->ev_start));
  BBX_CUDA_TRY(cudaEventCreate(&e->ev_stop));
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_upload, cudaEventDisableTiming));

  const uint32_t B = e->B, N = 2 * B;
  // twiddles exp(-2 pi i j / N), computed in double
  {
    std::vector<float2> tw(N);
    const double PI = 3.14159265358979323846264338327950288;
    for (uint32_t j = 0; j < N; j++) {
      double a = -2.0 * PI * (double)j / (double)N;
      tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->tw, sizeof(float2) * N));
    BBX_CUDA_TRY(cudaMemcpy(e->tw, tw.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
  }
  size_t xin_bytes = sizeof(float) * (size_t)e->n_in * e->xstride;
  for (int i = 0; i < 2; i++) {
    BBX_CUDA_TRY(cudaMalloc((void**)&e->xin[i], xin_bytes));
    BBX_CUDA_TRY(cudaMemset(e->xin[i], 0, xin_bytes));
  }
  size_t fdl_bytes = sizeof(float2) * (size_t)e->n_in * e->R * B;
  BBX_CUDA_TRY(cudaMalloc((void**)&e->fdl, fdl_bytes));
  BBX_CUDA_TRY(cudaMemset(e->fdl, 0, fdl_bytes));
  size_t ybuf_bytes = sizeof(float) * (size_t)e->n_streams * e->Rd;
  BBX_CUDA_TRY(cudaMalloc((void**)&e->ybuf, ybuf_bytes));
  BBX_CUDA_TRY(cudaMemset(e->ybuf, 0, ybuf_bytes));

  // plan capacities
  e->max_ctas = kNumSMs * 4;  // capacity for every tuning; the plan uses kNumSMs * mac_occ of them
  apply_tuning(e, cfg->mac_ctas_per_sm, cfg->mac_l2_keep_16ths, cfg->mac_time_tile);
  uint32_t terms = (e->mode == BBX_MODE_MIMO) ? e->n_paths : e->n_paths;
  e->max_jobs = 2 * e->n_streams + 1;
  e->max_segs = e->max_ctas + 2 * terms + 8;
  e->max_slots = e->max_ctas + e->max_jobs + 8;
  if ((rc = alloc_plan(e, e->plan_first))) return rc;
  if ((rc = alloc_plan(e, e->plan_steady))) return rc;
  size_t ypart_bytes = sizeof(float2) * (size_t)e->Tmax * e->max_slots * B;
  BBX_CUDA_TRY(cudaMalloc((void**)&e->ypart, ypart_bytes));
  BBX_CUDA_TRY(cudaMemset(e->ypart, 0, ypart_bytes));
  BBX_CUDA_TRY(cudaMalloc((void**)&e->nyq_part, sizeof(float) * (size_t)e->Tmax * e->max_slots));
  BBX_CUDA_TRY(cudaMemset(e->nyq_part, 0, sizeof(float) * (size_t)e->Tmax * e->max_slots));

  if (e->sh_world > 1) {
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_send, sizeof(float2) * (size_t)e->n_out * e->Tmax * B));
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_recv, sizeof(float2) * (size_t)e->sh_nloc * e->Tmax * B));
    std::vector<uint32_t> view(3 * (size_t)e->n_out);
    for (uint32_t o = 0; o < e->n_out; o++) {
      view[o] = (o >= e->sh_o0 && o < e->sh_o0 + e->sh_nloc) ? o - e->sh_o0 : 0u;
      view[e->n_out + o] = 1;
      view[2 * (size_t)e->n_out + o] = kNoJob;
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_view, sizeof(uint32_t) * view.size()));
    BBX_CUDA_TRY(cudaMemcpy(e->sh_view, view.data(), sizeof(uint32_t) * view.size(), cudaMemcpyHostToDevice));
  }

  // MIMO: tensor-core operands (time-batched calls only)
  if (e->mode == BBX_MODE_MIMO && cfg->mimo_tensor != 1 && e->Tmax >= e->tc_min_blocks && e->B >= 64) {
    uint32_t P2 = 1;
    while (P2 < e->Pmax) P2 <<= 1;
    // longer sums than kTcMaxK complex terms stay on the exact-fp32 SIMT MAC (accumulator truncation, mimo_tc.cuh)
    if ((uint64_t)e->n_in * P2 <= kTcMaxK) {
      if ((rc = tc_alloc(e))) return rc;
    }
  }

  // route tables
  {
    size_t off = 0;
    auto take = [&](size_t bytes) {
      size_t o = off;
      off += (bytes + 255) & ~(size_t)255;
      return o;
    };
    uint32_t ns = e->n_streams;
    e->roff_first = take(sizeof(uint32_t) * (e->n_out + 1));
    e->roff_stream = take(sizeof(uint32_t) * ns);
    e->roff_gain = take(sizeof(float) * ns);
    e->roff_dcur = take(sizeof(double) * ns);
    e->roff_dold = take(sizeof(double) * ns);
    e->roff_flags = take(sizeof(uint32_t) * ns);
    e->roff_icur = take(sizeof(uint32_t) * ns);
    e->roff_iold = take(sizeof(uint32_t) * ns);
    e->roff_entry = take(sizeof(RouteEntry) * ns);
    e->route_bytes = off;
    {
      int src = staging_alloc(e->route_stg, off);
      if (src) return src;
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->d_route, off));
  }
  BBX_CUDA_TRY(cudaDeviceSynchronize());
  *out = e;
  return BBX_OK;
}

int bbx_engine_destroy(bbx_engine* e) {
  if (!e) return BBX_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->s_in) cudaStreamSynchronize(e->s_in);
  if (e->s_out) cudaStreamSynchronize(e->s_out);
  if (e->s_aux) cudaStreamSynchronize(e->s_aux);
  cudaFree(e->tw);
  cudaFree(e->xin[0]);
  cudaFree(e->xin[1]);
  cudaFree(e->fdl);
  cudaFree(e->ypart);
  cudaFree(e->nyq_part);
  cudaFree(e->ybuf);
  for (int i = 0; i < 2; i++) {
    cudaFree(e->d_in[i]);
    cudaFree(e->d_out[i]);
    if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]);
    if (e->ev_comp[i]) cudaEventDestroy(e->ev_comp[i]);
    if (e->ev_d2h[i]) cudaEventDestroy(e->ev_d2h[i]);
  }
  if (e->ev_join_in) cudaEventDestroy(e->ev_join_in);
  if (e->ev_join_out) cudaEventDestroy(e->ev_join_out);
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  if (e->s_aux) cudaStreamDestroy(e->s_aux);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  cudaFree(e->flush_buf);
  for (void* pp : e->px_peer)
    if (pp) cudaIpcCloseMemHandle(pp);
  cudaFree(e->px_mem);
  cudaFree(e->px_done);
  cudaFree(e->px_view);
  if (e->px_status_h) cudaFreeHost(e->px_status_h);
  cudaFree(e->sh_send);
  cudaFree(e->sh_recv);
  cudaFree(e->sh_view);
  cudaFree(e->tc_hpack);
  cudaFree(e->tc_xb);
  cudaFree(e->tc_ftab_d);
  cudaFree(e->tc_fparts_d);
  cudaFree(e->tc_view);
  if (e->tc_status_h) cudaFreeHost(e->tc_status_h);
  cudaFree(e->tc_trace);
  if (e->tc_ftab_h) cudaFreeHost(e->tc_ftab_h);
  if (e->tc_fparts_h) cudaFreeHost(e->tc_fparts_h);
  cudaFree(e->d_route);
  staging_free(e->route_stg);
  for (MacPlan* pl : {&e->plan_first, &e->plan_steady}) {
    cudaFree(pl->d_blob);
    staging_free(pl->stg);
  }
  for (cudaEvent_t ev : e->mac_events) cudaEventDestroy(ev);
  if (e->ev_start) cudaEventDestroy(e->ev_start);
  if (e->ev_stop) cudaEventDestroy(e->ev_stop);
  if (e->ev_upload) cudaEventDestroy(e->ev_upload);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return BBX_OK;
}

uint32_t bbx_engine_get_ring_length(const bbx_engine* e) { return e ? e->Rd : 0; }
void* bbx_engine_get_stream(const bbx_engine* e) { return e ? (void*)e->stream : nullptr; }

int bbx_filter_create(bbx_engine* e, const float* ir, uint32_t length, bbx_filter** out) {
  BBX_REQUIRE(e && out, "bbx_filter_create: null argument");
  BBX_REQUIRE(ir || length == 0, "bbx_filter_create: null impulse response");
  const uint32_t B = e->B, N = 2 * B;
  uint32_t P = std::max(1u, ceil_div(length, B));
  BBX_REQUIRE(P <= e->Pmax, "impulse response of %u taps needs %u partitions, engine max_partitions is %u", length, P, e->Pmax);
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  // zero-padded windows [h[pB .. pB+B-1], 0^B]
  std::vector<float> pad((size_t)P * N, 0.0f);
  for (uint32_t i = 0; i < length; i++) pad[(size_t)(i / B) * N + (i % B)] = ir[i];
  float* d_pad = nullptr;
  BBX_CUDA_TRY(cudaMalloc((void**)&d_pad, sizeof(float) * pad.size()));
  bbx_filter* f = new bbx_filter();
  f->engine = e;
  f->P = P;
  f->H = nullptr;
  BBX_CUDA_TRY(cudaMalloc((void**)&f->H, sizeof(float2) * (size_t)P * B));
  BBX_CUDA_TRY(cudaMemcpyAsync(d_pad, pad.data(), sizeof(float) * pad.size(), cudaMemcpyHostToDevice, e->stream));
  // H = R2C(window) / N : the only normalisation of the whole path, exact (power of two)
  int rc = launch_rfft(B, d_pad, 0, N, f->H, 0, P, 0, e->tw, 1.0f / (float)N, 1, P, e->stream);
  e->launches++;
  cudaError_t se = cudaStreamSynchronize(e->stream);
  cudaFree(d_pad);
  if (rc || se != cudaSuccess) {
    if (!rc) set_error("filter transform failed: %s", cudaGetErrorString(se));
    cudaFree(f->H);
    delete f;
    return rc ? rc : BBX_ERR_CUDA;
  }
  *out = f;
  return BBX_OK;
}

int bbx_filter_destroy(bbx_filter* f) {
  if (!f) return BBX_OK;
  cudaSetDevice(f->engine->device);
  cudaStreamSynchronize(f->engine->stream);
  cudaFree(f->H);
  delete f;
  return BBX_OK;
}

uint32_t bbx_filter_partitions(const bbx_filter* f) { return f ? f->P : 0; }

int bbx_set_route(bbx_engine* e, uint32_t path, uint32_t input, uint32_t output, float gain) {
  BBX_REQUIRE(e != nullptr, "bbx_set_route: null engine");
  BBX_REQUIRE(e->mode == BBX_MODE_ROUTED, "bbx_set_route: engine is not in ROUTED mode");
  BBX_REQUIRE(path < e->n_paths && input < e->n_in && output < e->n_out, "bbx_set_route: index out of range");
  PathState& p = e->paths[path];
  if (p.input != input) e->steady_dirty = true;
  p.input = input;
  p.output = output;
  p.gain = gain;
  e->route_dirty = true;
  return BBX_OK;
}

int bbx_set_filter(bbx_engine* e, uint32_t path, const bbx_filter* filter, int crossfade, double delay) {
  BBX_REQUIRE(e != nullptr, "bbx_set_filter: null engine");
  BBX_REQUIRE(path < e->n_paths, "bbx_set_filter: path %u out of range", path);
  BBX_REQUIRE(!filter || filter->engine == e, "bbx_set_filter: filter belongs to another engine");
  BBX_REQUIRE(delay >= 0.0 && delay <= (double)e->cfg.max_delay, "bbx_set_filter: delay %g outside [0, max_delay=%u]", delay,
              e->cfg.max_delay);
  BBX_REQUIRE(e->mode != BBX_MODE_MIMO || delay == 0.0, "bbx_set_filter: MIMO mode has no per-path delay");
  BBX_REQUIRE(!(e->sh_world > 1 || e->comm) || !crossfade,
              "bbx_set_filter: the input-sharded MIMO engine switches filters without crossfade");
  PathState& p = e->paths[path];
  p.pend = filter;
  p.pend_delay = delay;
  p.xfade = crossfade != 0;
  p.has_pending = true;
  return BBX_OK;
}

// launch KERNEL<FMT, ACC> for a runtime (format, big-endian, typed-access) triple; ACC as in formats.cuh
#define BBX_PCM_LAUNCH_ACC(KERNEL, FMT, be, fast, GRID, STREAM, ARGS)            \
  do {                                                                            \
    if ((fast) && FMT != FMT_24) KERNEL<FMT, 2><<<GRID, 256, 0, STREAM>>>(ARGS);  \
    else if (be) KERNEL<FMT, 1><<<GRID, 256, 0, STREAM>>>(ARGS);                  \
    else KERNEL<FMT, 0><<<GRID, 256, 0, STREAM>>>(ARGS);                          \
  } while (0)
#define BBX_PCM_LAUNCH(KERNEL, fmt, be, fast, GRID, STREAM, ARGS)                          \
  do {                                                                                      \
    switch (fmt) {                                                                          \
      case FMT_16: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_16, be, fast, GRID, STREAM, ARGS); break;   \
      case FMT_24: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_24, be, fast, GRID, STREAM, ARGS); break;   \
      case FMT_32: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_32, be, fast, GRID, STREAM, ARGS); break;   \
      case FMT_F32: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_F32, be, fast, GRID, STREAM, ARGS); break; \
      default: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_F64, be, fast, GRID, STREAM, ARGS); break;      \
    }                                                                                       \
  } while (0)

int bbx_process_dev(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt,
                    int out_be, uint32_t out_channels, uint32_t nframes) {
  BBX_REQUIRE(e && in && out, "bbx_process: null argument");
  BBX_REQUIRE(infmt > FMT_UNKNOWN && infmt < FMT_COUNT && outfmt > FMT_UNKNOWN && outfmt < FMT_COUNT, "bbx_process: bad format");
  BBX_REQUIRE(nframes > 0 && nframes % e->B == 0, "bbx_process: nframes %u is not a positive multiple of the block size %u",
              nframes, e->B);
  const uint32_t B = e->B, T = nframes / B;
  BBX_REQUIRE(T <= e->Tmax, "bbx_process: %u blocks exceed max_blocks %u", T, e->Tmax);
  BBX_REQUIRE(in_channels >= e->n_in && out_channels >= e->n_out_pcm, "bbx_process: too few channels in the PCM buffers");
  BBX_REQUIRE(e->sh_world <= 1 || e->comm || e->px_on,
              "bbx_process: the input-sharded MIMO engine needs bbx_engine_set_comm() or bbx_engine_peer_attach()");
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  int rc;
  e->last_infmt = infmt;
  e->last_outfmt = outfmt;

  // ---- latch pending switches: they apply at this call's first block boundary ----
  bool any_pending = false, any_xfade = false;
  for (auto& p : e->paths)
    if (p.has_pending) {
      any_pending = true;
      if (p.xfade) any_xfade = true;
      else {  // hard switch: filter and delay jump now
        if (p.cur != p.pend) e->steady_dirty = true;
        p.cur = p.pend;
        p.delay = p.pend_delay;
        p.has_pending = false;
        e->route_dirty = true;
      }
    }
  uint32_t n_first = 0;
  std::vector<std::vector<JobTerm>> jobs;
  if (any_xfade) {
    // transitional plan for block 0: old filters as the main jobs, new filters as extra jobs
    make_jobs(e, false, jobs);
    std::vector<uint32_t> xjob(e->n_streams, kNoJob);
    if (e->mode == BBX_MODE_MIMO) {
      for (uint32_t o = 0; o < e->n_out; o++) {
        bool sw = false, differs = false;
        std::vector<JobTerm> nj;
        for (uint32_t i = 0; i < e->n_in; i++) {
          const PathState& p = e->paths[(size_t)o * e->n_in + i];
          const bbx_filter* f = p.has_pending ? p.pend : p.cur;
          if (p.has_pending) {
            sw = true;
            if (p.pend != p.cur) differs = true;
          }
          if (f) nj.push_back({f, i});
        }
        if (sw) {
          if (differs) {
            xjob[o] = (uint32_t)jobs.size();
            jobs.push_back(nj);
          } else {
            xjob[o] = kSameJob;
          }
        }
      }
    } else {
      for (uint32_t k = 0; k < e->n_paths; k++) {
        const PathState& p = e->paths[k];
        if (!p.has_pending) continue;
        if (p.pend == p.cur) xjob[k] = kSameJob;
        else {
          xjob[k] = (uint32_t)jobs.size();
          std::vector<JobTerm> nj;
          if (p.pend) nj.push_back({p.pend, p.input});
          jobs.push_back(nj);
        }
      }
    }
    if ((rc = build_plan(e, e->plan_first, jobs, xjob))) return rc;
    n_first = 1;
    if ((rc = upload_routes(e, true))) return rc;
    e->route_dirty = true;  // the steady-state tables follow after this call
    // commit the crossfaded switches
    for (auto& p : e->paths)
      if (p.has_pending) {
        if (p.cur != p.pend) e->steady_dirty = true;
        p.cur = p.pend;
        p.delay = p.pend_delay;
        p.has_pending = false;
      }
  } else if (e->route_dirty) {
    if ((rc = upload_routes(e, false))) return rc;
    e->route_dirty = false;
  }
  (void)any_pending;
  if (e->steady_dirty || !e->plan_steady.valid) {
    make_jobs(e, false, jobs);
    if ((rc = build_plan(e, e->plan_steady, jobs, std::vector<uint32_t>()))) return rc;
    e->steady_dirty = false;
    e->tc_dirty = true;
  }
  // MIMO calls of >= tc_min_blocks blocks without a crossfade in flight run the per-bin GEMM on the tensor cores
  const bool use_tc = e->tc_on && n_first == 0 && T >= e->tc_min_blocks && e->plan_steady.total_rows > 0;
  if (use_tc && e->tc_dirty) {
    if ((rc = tc_pack_filters(e))) return rc;
  }

  // ---- 1. PCM -> planar fp32 ----
  {
    PcmInArgs a;
    a.pcm = (const uint8_t*)in;
    a.fmt = infmt;
    a.be = in_be;
    a.in_channels = in_channels;
    a.n_inputs = e->n_in;
    a.B = B;
    a.T = T;
    a.xin_cur = e->xin[e->parity];
    a.xin_prev = e->xin[e->parity ^ 1];
    a.xstride = e->xstride;
    a.prev_off = e->tprev * B;
    {
      const uint32_t bps = fmt_bytes(infmt);
      a.fast = (!in_be && bps != 3 && ((uintptr_t)in % bps) == 0) ? 1 : 0;  // frame stride = in_channels * bps is aligned too
    }
    // wide tiles pay when the channel axis fills the lanes; few-channel engines keep the finer grid
    if (B % 128 == 0 && e->n_in >= 16)
      BBX_PCM_LAUNCH(k_pcm_in128, infmt, in_be, a.fast, dim3((T + 1) * B / 128, ceil_div(e->n_in, 32)), st, a);
    else BBX_PCM_LAUNCH(k_pcm_in, infmt, in_be, a.fast, dim3((T + 1) * B / 32, ceil_div(e->n_in, 32)), st, a);
    BBX_CUDA_TRY(cudaGetLastError());
    e->launches++;
  }
  // ---- 2. forward transforms into the FDL ----
  if ((rc = launch_rfft(B, e->xin[e->parity], e->xstride, B, e->fdl, (uint64_t)e->R * B, e->R, e->head, e->tw, 1.0f, e->n_in, T, st)))
    return rc;
  e->launches++;
  // ---- 3. FDL multiply-accumulate ----
  if (use_tc) {
    if ((rc = launch_mimo_tc(e, T))) return rc;
  } else if (n_first) {
    if ((rc = launch_mac(e, e->plan_first, 0, 1))) return rc;
    if ((rc = launch_mac(e, e->plan_steady, 1, T - 1))) return rc;
  } else {
    if ((rc = launch_mac(e, e->plan_steady, 0, T))) return rc;
  }
  // ---- 3b. input-sharded MIMO: sum the partial spectra over the ranks, keep the local outputs ----
  if (e->sh_world > 1 || e->comm) {
    const PlanView pv = use_tc ? tc_plan_view(e) : e->plan_steady.view();
    if (e->px_on) {
      e->px_epoch++;
      const uint32_t parity = e->px_epoch & 1u;
      k_gather_spectra_peer<<<dim3(e->n_out, T), 256, 0, st>>>(e->ypart, use_tc ? nullptr : e->nyq_part, e->max_slots, pv,
                                                               e->px_table, e->sh_world, e->sh_rank, e->sh_nloc, B, T, e->px_half,
                                                               parity, e->px_epoch, e->px_done);
      BBX_CUDA_TRY(cudaGetLastError());
      k_peer_wait<<<1, 32, 0, st>>>((const uint32_t*)(e->px_mem + e->px_flag_off), e->sh_world, parity, e->px_epoch, e->px_status);
      BBX_CUDA_TRY(cudaGetLastError());
      e->launches += 2;
    } else {
      k_gather_spectra<<<dim3(e->n_out, T), 256, 0, st>>>(e->ypart, use_tc ? nullptr : e->nyq_part, e->max_slots, pv, e->sh_send, B, T);
      BBX_CUDA_TRY(cudaGetLastError());
      e->launches++;
      if ((rc = comm_reduce_scatter_f32(e->comm, (const float*)e->sh_send, (float*)e->sh_recv, (size_t)e->sh_nloc * T * B * 2, st)))
        return rc;
    }
  }
  // ---- 4. inverse transforms, crossfade, delay ring ----
  if ((rc = launch_irfft(e, T, n_first, use_tc, st))) return rc;
  e->launches++;
  // ---- 5. delay read, mixdown, output format ----
  {
    PcmOutArgs a;
    const uint32_t obps = fmt_bytes(outfmt);
    a.pcm = (uint8_t*)out;
    a.fmt = outfmt;
    a.be = out_be;
    a.out_channels = out_channels;
    a.n_outputs = e->n_out_pcm;
    a.B = B;
    a.T = T;
    a.ybuf = e->ybuf;
    a.Rd = e->Rd;
    a.wpos0 = e->wpos;
    a.fractional = e->cfg.fractional_delay;
    a.fast = (!out_be && obps != 3 && ((uintptr_t)out % obps) == 0) ? 1 : 0;
    a.rv = route_view(e);
    // wide tiles for many outputs with integer delays (their ring reads batch); mixdowns of many paths into few
    // outputs and fractional delays (14-tap double-precision reads) keep the finer grid
    // mixdowns (at least four paths per output on average, few outputs): the (route, frame)-parallel kernel
    const bool mix = e->n_out_pcm <= 32 && e->n_routes_pcm <= kMixMaxRoutes && e->n_routes_pcm >= 4 * e->n_out_pcm &&
                     e->pcm_out_mix;
    if (mix) BBX_PCM_LAUNCH(k_pcm_out_mix, outfmt, out_be, a.fast, dim3(T * B / 32), st, a);
    else if (B % 128 == 0 && e->n_out_pcm >= 16 && !e->cfg.fractional_delay)
      BBX_PCM_LAUNCH(k_pcm_out128, outfmt, out_be, a.fast, dim3(T * B / 128, ceil_div(e->n_out_pcm, 32)), st, a);
    else BBX_PCM_LAUNCH(k_pcm_out, outfmt, out_be, a.fast, dim3(T * B / 32, ceil_div(e->n_out_pcm, 32)), st, a);
    BBX_CUDA_TRY(cudaGetLastError());
    e->launches++;
  }
  // ---- advance the state ----
  e->head = (e->head + T) % e->R;
  e->wpos = (e->wpos + T * B) % e->Rd;
  e->parity ^= 1;
  e->tprev = T;
  return BBX_OK;
}

// Device-side address of a pinned host buffer (cudaHostAlloc / cudaHostRegister under unified addressing), or false.
static bool mapped_device_ptr(const void* host, void** dev) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
  *dev = at.devicePointer;
  return true;
}

int bbx_process_async(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt,
                      int out_be, uint32_t out_channels, uint32_t nframes) {
  BBX_REQUIRE(e && in && out, "bbx_process: null argument");
  BBX_REQUIRE(infmt > FMT_UNKNOWN && infmt < FMT_COUNT && outfmt > FMT_UNKNOWN && outfmt < FMT_COUNT, "bbx_process: bad format");
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  const uint32_t ibps = fmt_bytes(infmt), obps = fmt_bytes(outfmt);
  size_t in_bytes = (size_t)nframes * in_channels * ibps;
  size_t out_bytes = (size_t)nframes * out_channels * obps;
  size_t need = std::max(in_bytes, out_bytes);
  // Latency path (real-time callers: one or a few blocks per call): a short buffer in pinned host memory that the
  // device can address is read by k_pcm_in / written by k_pcm_out straight over PCIe -- no copy-engine operation and no
  // cross-stream hand-off on that side.  Decided per side, and only where the kernel's accesses suit the bus: typed
  // little-endian samples (no 3-byte formats) and at least 128 contiguous bytes of used channels per frame (a warp
  // covers 32 channels of one frame); narrow or byte-wise layouts would turn into many small PCIe transactions and
  // stay on the staged path, like pageable buffers.
  void *din = nullptr, *dout = nullptr;
  const bool typed_in = !in_be && ibps != 3 && ((uintptr_t)in % ibps) == 0;
  const bool typed_out = !out_be && obps != 3 && ((uintptr_t)out % obps) == 0;
  const bool direct_in = in_bytes <= e->direct_io_max_bytes && typed_in && (size_t)e->n_in * ibps >= 128 && mapped_device_ptr(in, &din);
  const bool direct_out =
      out_bytes <= e->direct_io_max_bytes && typed_out && (size_t)e->n_out_pcm * obps >= 128 && mapped_device_ptr(out, &dout);
  if (direct_in && direct_out) {
    e->direct_calls++;
    return bbx_process_dev(e, din, infmt, in_be, in_channels, dout, outfmt, out_be, out_channels, nframes);
  }
  if (!e->d_in[0] || e->d_io_bytes < need) {
    BBX_CUDA_TRY(cudaStreamSynchronize(e->s_in));
    BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
    BBX_CUDA_TRY(cudaStreamSynchronize(e->s_out));
    e->d_io_bytes = std::max(need, e->d_io_bytes);
    for (int i = 0; i < 2; i++) {
      cudaFree(e->d_in[i]);
      cudaFree(e->d_out[i]);
      e->d_in[i] = e->d_out[i] = nullptr;
      BBX_CUDA_TRY(cudaMalloc((void**)&e->d_in[i], e->d_io_bytes));
      BBX_CUDA_TRY(cudaMalloc((void**)&e->d_out[i], e->d_io_bytes));
    }
  }
  const int k = (int)(e->host_calls & 1);
  e->host_calls++;
  if (direct_in || direct_out) e->direct_calls++;
  // H2D on the input-copy stream, once the kernels of call n-2 have finished reading this staging buffer
  bool fed = false;
  if (!direct_in) {
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_in, e->ev_comp[k], 0));
    BBX_CUDA_TRY(cudaMemcpyAsync(e->d_in[k], in, in_bytes, cudaMemcpyHostToDevice, e->s_in));
    fed = true;
  }
  if (!direct_out && out_channels > e->n_out_pcm) {
    // channels beyond n_outputs keep the caller's bytes: seed the output staging with them
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_in, e->ev_d2h[k], 0));
    BBX_CUDA_TRY(cudaMemcpyAsync(e->d_out[k], out, out_bytes, cudaMemcpyHostToDevice, e->s_in));
    fed = true;
  }
  if (fed) {
    BBX_CUDA_TRY(cudaEventRecord(e->ev_h2d[k], e->s_in));
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_h2d[k], 0));
  }
  // kernels on the engine stream
  if (!direct_out) BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_d2h[k], 0));
  int rc = bbx_process_dev(e, direct_in ? din : e->d_in[k], infmt, in_be, in_channels, direct_out ? dout : e->d_out[k], outfmt, out_be,
                           out_channels, nframes);
  if (rc) return rc;
  BBX_CUDA_TRY(cudaEventRecord(e->ev_comp[k], e->stream));
  if (!direct_out) {
    // D2H on the output-copy stream
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_out, e->ev_comp[k], 0));
    BBX_CUDA_TRY(cudaMemcpyAsync(out, e->d_out[k], out_bytes, cudaMemcpyDeviceToHost, e->s_out));
    BBX_CUDA_TRY(cudaEventRecord(e->ev_d2h[k], e->s_out));
  }
  return BBX_OK;
}

int bbx_engine_sync(bbx_engine* e) {
  BBX_REQUIRE(e != nullptr, "bbx_engine_sync: null engine");
  BBX_CUDA_TRY(cudaStreamSynchronize(e->s_in));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->s_out));
  if (e->px_status_h && *(volatile int*)e->px_status_h) {
    set_error("peer mixdown: rank %d never published its partial spectra (timed out); results are invalid",
              *(volatile int*)e->px_status_h - 1);
    return BBX_ERR_CUDA;
  }
  if (e->tc_status_h && *(volatile int*)e->tc_status_h) {
    // fail loudly: the output of that call is not valid
    set_error("k_mimo_tc: a barrier wait timed out inside the tensor-core kernel (status %d); results are invalid",
              *(volatile int*)e->tc_status_h);
    return BBX_ERR_CUDA;
  }
  return BBX_OK;
}

int bbx_process(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt,
                int out_be, uint32_t out_channels, uint32_t nframes) {
  int rc = bbx_process_async(e, in, infmt, in_be, in_channels, out, outfmt, out_be, out_channels, nframes);
  if (rc) return rc;
  return bbx_engine_sync(e);
}

int bbx_blockconvolver_convolve(bbx_engine* e, const float* in, float* out) {
  BBX_REQUIRE(e && in && out, "bbx_blockconvolver_convolve: null argument");
  BBX_REQUIRE(e->n_in == 1 && e->n_out == 1, "bbx_blockconvolver_convolve: engine must be single-channel");
  return bbx_process(e, in, FMT_F32, 0, 1, out, FMT_F32, 0, 1, e->B);
}

int bbx_engine_timer_start(bbx_engine* e) {
  BBX_REQUIRE(e != nullptr, "null engine");
  // nothing of the timed region may start before the start event: drain, record, and fence the other streams
  int rc = bbx_engine_sync(e);
  if (rc) return rc;
  BBX_CUDA_TRY(cudaEventRecord(e->ev_start, e->s_in));
  BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_start, 0));
  BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_out, e->ev_start, 0));
  return BBX_OK;
}
int bbx_engine_timer_stop(bbx_engine* e, float* elapsed_ms) {
  BBX_REQUIRE(e && elapsed_ms, "null argument");
  // the stop event follows everything enqueued on the copy streams and the engine stream
  BBX_CUDA_TRY(cudaEventRecord(e->ev_join_in, e->s_in));
  BBX_CUDA_TRY(cudaEventRecord(e->ev_join_out, e->s_out));
  BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join_in, 0));
  BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join_out, 0));
  BBX_CUDA_TRY(cudaEventRecord(e->ev_stop, e->stream));
  BBX_CUDA_TRY(cudaEventSynchronize(e->ev_stop));
  BBX_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, e->ev_start, e->ev_stop));
  return BBX_OK;
}
uint64_t bbx_engine_launch_count(const bbx_engine* e) { return e ? e->launches : 0; }

int bbx_engine_profile_mac(bbx_engine* e, int enable) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->profile_mac = enable != 0;
  e->mac_events_used = 0;
  e->mac_ms_total = 0.0;
  e->mac_launches = e->mac_units = e->mac_bytes = 0;
  return BBX_OK;
}

int bbx_engine_mac_time(bbx_engine* e, float* total_ms, uint64_t* launches, uint64_t* channel_blocks,
                        uint64_t* algorithmic_bytes) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (size_t i = 0; i + 1 < e->mac_events_used; i += 2) {
    float ms = 0.f;
    BBX_CUDA_TRY(cudaEventElapsedTime(&ms, e->mac_events[i], e->mac_events[i + 1]));
    e->mac_ms_total += ms;
  }
  e->mac_events_used = 0;
  if (total_ms) *total_ms = (float)e->mac_ms_total;
  if (launches) *launches = e->mac_launches;
  if (channel_blocks) *channel_blocks = e->mac_units;
  if (algorithmic_bytes) *algorithmic_bytes = e->mac_bytes;
  return BBX_OK;
}

int bbx_engine_set_tuning(bbx_engine* e, uint32_t ctas_per_sm, uint32_t l2_keep_16ths, uint32_t time_tile) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  apply_tuning(e, ctas_per_sm, l2_keep_16ths, time_tile);
  return BBX_OK;
}

int bbx_engine_set_mixdown_kernel(bbx_engine* e, int per_output) {
  BBX_REQUIRE(e != nullptr, "bbx_engine_set_mixdown_kernel: null engine");
  e->pcm_out_mix = per_output == 0;
  return BBX_OK;
}

int bbx_engine_set_direct_io(bbx_engine* e, size_t max_bytes) {
  BBX_REQUIRE(e != nullptr, "bbx_engine_set_direct_io: null engine");
  e->direct_io_max_bytes = max_bytes;
  return BBX_OK;
}
uint64_t bbx_engine_direct_calls(const bbx_engine* e) { return e ? e->direct_calls : 0; }

int bbx_engine_set_comm(bbx_engine* e, bbx_comm* c) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_REQUIRE(e->mode == BBX_MODE_MIMO, "bbx_engine_set_comm: only the MIMO engine has a collective");
  BBX_REQUIRE(!c || ((uint32_t)comm_world(c) == e->sh_world && (uint32_t)comm_rank(c) == e->sh_rank),
              "bbx_engine_set_comm: communicator (rank %d of %d) does not match the engine (rank %u of %u)", comm_rank(c),
              comm_world(c), e->sh_rank, e->sh_world);
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (c && !e->sh_send) {
    // world == 1 with a communicator: the sharded code path on one GPU (tests)
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_send, sizeof(float2) * (size_t)e->n_out * e->Tmax * e->B));
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_recv, sizeof(float2) * (size_t)e->sh_nloc * e->Tmax * e->B));
    std::vector<uint32_t> view(3 * (size_t)e->n_out);
    for (uint32_t o = 0; o < e->n_out; o++) {
      view[o] = o;
      view[e->n_out + o] = 1;
      view[2 * (size_t)e->n_out + o] = kNoJob;
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_view, sizeof(uint32_t) * view.size()));
    BBX_CUDA_TRY(cudaMemcpy(e->sh_view, view.data(), sizeof(uint32_t) * view.size(), cudaMemcpyHostToDevice));
  }
  e->comm = c;
  return BBX_OK;
}

int bbx_engine_peer_export(bbx_engine* e, uint8_t* handle64) {
  BBX_REQUIRE(e && handle64, "bbx_engine_peer_export: null argument");
  BBX_REQUIRE(e->mode == BBX_MODE_MIMO && e->sh_world > 1 && e->sh_world <= 16,
              "bbx_engine_peer_export: only the input-sharded MIMO engine (2..16 ranks) exchanges spectra");
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  if (!e->px_mem) {
    e->px_half = (size_t)e->sh_nloc * e->sh_world * e->Tmax * e->B;
    e->px_flag_off = (2 * e->px_half * sizeof(float2) + 255) & ~(size_t)255;
    const size_t bytes = e->px_flag_off + 2 * sizeof(uint32_t) * e->sh_world;
    BBX_CUDA_TRY(cudaMalloc((void**)&e->px_mem, bytes));
    BBX_CUDA_TRY(cudaMemset(e->px_mem, 0, bytes));
    BBX_CUDA_TRY(cudaMalloc((void**)&e->px_done, sizeof(uint32_t)));
    BBX_CUDA_TRY(cudaMemset(e->px_done, 0, sizeof(uint32_t)));
    std::vector<uint32_t> view(3 * (size_t)e->n_out);
    for (uint32_t o = 0; o < e->n_out; o++) {
      const bool mine = o >= e->sh_o0 && o < e->sh_o0 + e->sh_nloc;
      view[o] = mine ? (o - e->sh_o0) * e->sh_world : 0u;
      view[e->n_out + o] = e->sh_world;
      view[2 * (size_t)e->n_out + o] = kNoJob;
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->px_view, sizeof(uint32_t) * view.size()));
    BBX_CUDA_TRY(cudaMemcpy(e->px_view, view.data(), sizeof(uint32_t) * view.size(), cudaMemcpyHostToDevice));
    BBX_CUDA_TRY(cudaHostAlloc((void**)&e->px_status_h, sizeof(int), cudaHostAllocMapped));
    *e->px_status_h = 0;
    BBX_CUDA_TRY(cudaHostGetDevicePointer((void**)&e->px_status, e->px_status_h, 0));
    BBX_CUDA_TRY(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  BBX_CUDA_TRY(cudaIpcGetMemHandle(&h, e->px_mem));
  static_assert(sizeof(h) == BBX_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  memcpy(handle64, &h, sizeof(h));
  return BBX_OK;
}

int bbx_engine_peer_attach(bbx_engine* e, const uint8_t* handles) {
  BBX_REQUIRE(e && handles, "bbx_engine_peer_attach: null argument");
  BBX_REQUIRE(e->px_mem != nullptr, "bbx_engine_peer_attach: call bbx_engine_peer_export first");
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (uint32_t r = 0; r < e->sh_world; r++) {
    uint8_t* base = e->px_mem;
    if (r != e->sh_rank) {
      if (!e->px_peer[r]) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * BBX_PEER_HANDLE_BYTES, sizeof(h));
        BBX_CUDA_TRY(cudaIpcOpenMemHandle(&e->px_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
      }
      base = (uint8_t*)e->px_peer[r];
    }
    e->px_table.data[r] = (float2*)base;
    e->px_table.flags[r] = (uint32_t*)(base + e->px_flag_off);
  }
  e->px_on = true;
  return BBX_OK;
}

int bbx_engine_tensor_status(bbx_engine* e, uint64_t* launches, int* status) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (launches) *launches = e->tc_launches;
  if (status) {
    *status = 0;
    if (e->tc_status_h) *status = *(volatile int*)e->tc_status_h;
  }
  return BBX_OK;
}

int bbx_probe_fp32_tflops(int device, float seconds, float* burst, float* sustained) {
  BBX_REQUIRE(burst || sustained, "bbx_probe_fp32_tflops: null outputs");
  int rc = require_device();
  if (rc) return rc;
  BBX_CUDA_TRY(cudaSetDevice(device));
  const int grid = kNumSMs * 4, iters = 4096;
  const double flops = 8.0 * 16 * (double)iters * grid * 256;  // 16 complex MACs = 64 FMA = 128 flop per thread and iteration
  float2* out = nullptr;
  BBX_CUDA_TRY(cudaMalloc((void**)&out, sizeof(float2) * (size_t)grid * 256));
  cudaEvent_t e0, e1;
  BBX_CUDA_TRY(cudaEventCreate(&e0));
  BBX_CUDA_TRY(cudaEventCreate(&e1));
  const float2 h = make_float2(1.0001f, 0.0001f), x = make_float2(0.9999f, 0.0002f);
  k_fp32_probe<<<grid, 256>>>(out, 64, h, x);  // warm-up
  BBX_CUDA_TRY(cudaDeviceSynchronize());
  float best = 0.f;
  for (int r = 0; r < 5; r++) {  // burst: best of five isolated launches
    BBX_CUDA_TRY(cudaEventRecord(e0));
    k_fp32_probe<<<grid, 256>>>(out, iters, h, x);
    BBX_CUDA_TRY(cudaEventRecord(e1));
    BBX_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    BBX_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (ms > 0.f) best = std::max(best, (float)(flops / (ms * 1e-3) / 1e12));
  }
  if (burst) *burst = best;
  if (sustained) {  // back-to-back launches for `seconds`: the rate under the power cap
    const int n = std::max(1, (int)(seconds * 1e3f / 1.3f));
    BBX_CUDA_TRY(cudaEventRecord(e0));
    for (int r = 0; r < n; r++) k_fp32_probe<<<grid, 256>>>(out, iters, h, x);
    BBX_CUDA_TRY(cudaEventRecord(e1));
    BBX_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    BBX_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *sustained = ms > 0.f ? (float)(flops * n / (ms * 1e-3) / 1e12) : 0.f;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return BBX_OK;
}

int bbx_engine_tensor_trace(bbx_engine* e, uint64_t* out, uint32_t max_ctas) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaSetDevice(e->device));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (!out) {  // enable (max_ctas > 0) or disable
    cudaFree(e->tc_trace);
    e->tc_trace = nullptr;
    e->tc_trace_ctas = 0;
    if (max_ctas) {
      BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_trace, sizeof(unsigned long long) * 16 * (size_t)max_ctas));
      BBX_CUDA_TRY(cudaMemset(e->tc_trace, 0, sizeof(unsigned long long) * 16 * (size_t)max_ctas));
      e->tc_trace_ctas = max_ctas;
    }
    return BBX_OK;
  }
  BBX_REQUIRE(e->tc_trace && max_ctas <= e->tc_trace_ctas, "bbx_engine_tensor_trace: tracing is not enabled for that many CTAs");
  BBX_CUDA_TRY(cudaMemcpy(out, e->tc_trace, sizeof(unsigned long long) * 16 * (size_t)max_ctas, cudaMemcpyDeviceToHost));
  return BBX_OK;
}

int bbx_engine_flush_l2(bbx_engine* e, size_t bytes) {
  BBX_REQUIRE(e != nullptr, "null engine");
  bytes = (bytes + 15) & ~(size_t)15;
  if (e->flush_bytes < bytes) {
    cudaStreamSynchronize(e->stream);
    cudaFree(e->flush_buf);
    e->flush_buf = nullptr;
    BBX_CUDA_TRY(cudaMalloc((void**)&e->flush_buf, bytes));
    e->flush_bytes = bytes;
  }
  k_flush<<<kNumSMs * 8, 256, 0, e->stream>>>(e->flush_buf, bytes / 16);
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

}  // extern "C"
