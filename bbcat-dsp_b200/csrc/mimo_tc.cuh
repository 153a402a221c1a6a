// mimo_tc.cuh -- MIMO frequency-domain mixdown as a per-bin complex GEMM on the tcgen05 tensor cores.
//
// Y_o[k] = sum_i sum_p H_{o,i}[p][k] * FDL_i[(head - p)][k].  With T block-steps in one
// call this is, for every bin k, a dense complex GEMM
//       Y[o][t] = sum_j  H[o][j] * X[j][t],     j = (i, p),  X[(i,p)][t] = FDL_i[slot(t - p)]
// (M = n_out, K = n_in * P, N = T).  The complex product is embedded in a real one with PLANAR operands, so that one
// MMA is 128 rows x 2N columns and its K index is the complex index j itself:
//       A = [ Hr ]  (rows 0..63:  re of H for output o)        B = [ Xr | Xi ]  (columns 0..N-1: re of X at block-step t,
//           [ Hi ]  (rows 64..127: im of H for output o)                         columns N..2N-1: im)
//       D = A B = [ HrXr  HrXi ]        Y.re = HrXr - HiXi   (quadrants 00 and 11)
//                 [ HiXr  HiXi ]        Y.im = HrXi + HiXr   (quadrants 01 and 10), combined by the epilogue.
// Against the interleaved embedding of round 1 ([Hr, -Hi; Hi, Hr] x [Xr; Xi], 128 x N x 2K) the flops are the same, but
// every MMA is twice as long (N = 128: 64 cycles instead of 32 -- round 1's issue loop could not feed 32-cycle MMAs) and the
// A operand is half as large (every H value is split into TF32 parts once, not twice), which halves the producers' work.
// fp32 accuracy (SNR >= 110 dB) needs more than one TF32 pass: both operands are split v = hi + lo with
// hi = tf32_rna(v), and three MMAs (lo*hi, hi*lo, hi*hi) accumulate into fp32 TMEM tiles (see kTcAccTiles).
//
// Operand layouts in HBM (bin-major, written by the pack kernels below):
//   Hpack[og][k][chunk][re | im][g][64] float4
//                                 planar: the real (then the imaginary) parts of the four complex K elements
//                                 j = 16 chunk + 4g .. + 3 of one output per float4;  j = i*P2 + p',
//                                 p' = P2-1-p (partition order reversed so a column of B is a contiguous
//                                 run of the time axis), P2 = power of two >= P, K padded to 16 with zeros
//   Xq[k][column tile][chunk][hi | lo][input row of the chunk][seg] float2
//                                 the FDL runs of one chunk, already split into TF32 hi and lo parts and laid out
//                                 exactly as the kernel stages them: ONE contiguous TMA copy per chunk.  Element wl of
//                                 a row is block (tile start + p'0 + wl - (P2-1)) of this call (negative = history,
//                                 beyond the call = 0); seg = N + P2 (P2 < 16: 16/P2 rows per chunk) or N + 16 (one row)
// One CTA owns kTcBins adjacent bins (their 8-byte output writes fill one 32-byte sector in L2) and walks K in chunks
// of 16 complex (two MMA k-steps) in three decoupled pipelines (mbarriers only, no block-wide barrier in the main loop):
//   loader warp   TMA bulk copies (cp.async.bulk) of the chunk's raw H and FDL runs into a raw smem ring
//   4 x 4 producer warps (group g: chunks g mod 4)  raw H -> re (rows < 64) or im (rows >= 64) part -> hi/lo split -> A tile
//                 into TENSOR MEMORY (tcgen05.st); raw FDL runs (already split by k_mimo_pack_x) -> the Toeplitz B tile in
//                 shared memory (canonical no-swizzle K-major layout: a 16-byte row = four consecutive j of one column)
//   4 epilogue warps  accumulators (TMEM) -> sum of the three tiles -> quadrants through two smem tiles -> (re, im) -> HBM
//   2 MMA warps   one lane each issues tcgen05.mma (A from TMEM, B from smem): warp 20 the 2 hi*hi MMAs of the chunk,
//                 warp 21 the 4 small-term MMAs (independent accumulator tiles); their tcgen05.commit's release the stage
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bbx {

static constexpr int kTcGroups = 4;        // producer groups: group g converts chunks it = g (mod 4) into operand stage g, so
                                          // the fences, store waits and barrier round trips of four chunks overlap
static constexpr int kTcProducers = 128;  // threads per group (4 warps = the four TMEM lane quarters); warps 0..15 produce
static constexpr int kTcWarpEpi = 16;     // warps 16..19: accumulator read-out (warp & 3 = TMEM lane quarter)
static constexpr int kTcEpiThreads = 128;
static constexpr int kTcWarpMma = 20;     // warps 20, 21: one elected lane each issues the hi*hi MMAs / the small-term MMAs
static constexpr int kTcMmaWarps = 2;     //   (an N = 64 MMA is shorter than one lane's issue overhead: two issue streams)
static constexpr int kTcWarpLoad = 22;    // one elected lane issues the TMA bulk copies (a second loader warp changes nothing)
static constexpr int kTcLoadWarps = 1;
static constexpr int kTcThreads = 736;
static constexpr int kTcBins = 4;      // adjacent bins per CTA
static constexpr int kTcStages = 4;    // converted operand stages: A in TMEM, B in shared memory
static constexpr int kTcPrefetch = 16;    // chunks of packed H prefetched into L2 ahead of their copy
static constexpr int kTcRawStagesMax = 8; // raw operand stages (TMA bulk copies from HBM): as many as fit, even
static constexpr int kTcChunk = 16;    // complex K elements per stage = 16 tf32 = 2 MMA k-steps
static constexpr int kTcRows = 128;    // accumulator rows: 64 outputs x (Hr, Hi)
static constexpr int kTcNmax = 64;     // block-steps per CTA
static constexpr int kTcN2max = 2 * kTcNmax;  // MMA columns: (Xr | Xi) of every block-step
static constexpr uint32_t kTcBHalf = kTcN2max * kTcChunk * 4;             // 8 KB:  B_hi (then B_lo)
static constexpr uint32_t kTcStageBytes = 2 * kTcBHalf;                   // 16 KB of shared memory per stage
static constexpr uint32_t kTcTileBytes = kTcNmax * 64 * 16;               // 64 KB: epilogue tile [t][o][bin pair] float2
static_assert(kTcBins % 2 == 0, "the epilogue stores pairs of bins");
// raw stage: the FDL runs of the chunk: (16 / P2) input rows of
// N + P2 complex (P2 < 16) or one row of N + 16, hi and lo parts.  Its size depends on (P2, N): the host passes
// the stage size and the number of stages that fit (MimoTcArgs::raw_stage_bytes / raw_stages).
static constexpr uint32_t kTcOffTile = kTcStages * kTcStageBytes;
static constexpr uint32_t kTcOffBar = kTcOffTile + kTcTileBytes;
static constexpr uint32_t kTcOffRaw = kTcOffBar + 256;
static constexpr uint32_t kTcSmemMax = 227 * 1024;
// TMEM (512 columns x 128 lanes): columns 0..383 = three 128-column accumulator tiles, columns 384..511 = the A
// operand ring (per stage 16 columns of A_hi and 16 of A_lo: row = lane, K along columns).  A never touches
// shared memory: the producers write it with tcgen05.st, the MMA reads it from TMEM ([a-tmem] operand form), which
// takes the larger operand off the shared-memory pipe (the bottleneck of the SS form, profiles/).
// The tensor core truncates the fp32 accumulator on every MMA, a bias that grows with the number of sequential
// accumulations (one tile for everything: 109 dB SNR at K = 512 complex).  The hi*hi products therefore alternate
// between two tiles (the two k-steps of a chunk) and the small lo*hi / hi*lo terms have their own tile, so the dominant
// sums see K/16 accumulations (32 at K = 512); the epilogue adds the three tiles in fp32 round-to-nearest.
static constexpr int kTcAccTiles = 3;
static constexpr uint32_t kTcAccCols = kTcAccTiles * kTcN2max;            // 384
static constexpr uint32_t kTcAStageCols = 2 * kTcChunk;                   // 32: A_hi | A_lo
static constexpr uint32_t kTcTmemCols = 512;
static constexpr uint32_t kTcMaxK = 1024;                                 // complex K verified against the tolerance

struct MimoTcArgs {
  const float4* hpack;
  const float2* xb;
  float2* ypart;     // [t][slot_stride][B], slot = output
  int* status;       // mapped host memory: set non-zero when a barrier wait times out (never hang the device)
  uint32_t B, n_in, n_out, P2log, G /* float4 K groups = Kc/2 */, T, slot_stride;
  uint64_t xbin;     // float2 elements per bin of xb = column tiles * chunks * (2 * ninp * seg)
  uint32_t raw_stage_bytes, raw_stages;  // raw ring geometry (stages: even, <= kTcRawStagesMax)
  unsigned long long* trace;             // optional [gridDim.x][16] cycle counters of the warp roles (NULL = off)
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

// bounded wait: returns false after ~1e6 polls (never hang the device); one poll on the fast path
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 20); spin++)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  return mbar_wait_slow(bar, parity);
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

// A operand in tensor memory, B through a shared-memory descriptor
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 8 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void st_tmem8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

// 16 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void st_tmem16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
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

// a barrier wait gave up: tell the host (status lives in mapped pinned memory, read by bbx_engine_sync)
__device__ __forceinline__ void report_timeout(int* status, int code) {
  *reinterpret_cast<volatile int*>(status) = code;
  __threadfence_system();
}

// timed wait for the optional role trace (cycles spent waiting are added to *acc)
__device__ __forceinline__ bool mbar_wait_t(uint32_t bar, uint32_t parity, bool trace, unsigned long long& acc) {
  if (!trace) return mbar_wait(bar, parity);
  const long long t0 = clock64();
  const bool r = mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
  return r;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float4 ld_shared4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ float2 ld_shared2(uint32_t addr) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr) : "memory");
  return r;
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (16-byte aligned addresses and size)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// L2 prefetch of a run the loader will copy a few chunks later (16-byte aligned address and size)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_shared4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_shared2(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

}  // namespace tc

namespace tc {

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// B tile descriptors (SWIZZLE_NONE, K-major): high word = SBO 128 bytes | version 1, low word = LBO (N * 16 bytes)
// | address >> 4: advancing a descriptor is one 32-bit add on the low word
static constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
template <uint32_t N>
static constexpr uint32_t kDescLo = ((N * 16u) >> 4) << 16;

template <bool ACC>
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t desc_lo, uint32_t idesc) {
  const uint64_t db = ((uint64_t)kDescHi << 32) | desc_lo;
  if (ACC)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc)
        : "memory");
}

// the MMAs of one chunk (2 k-steps of 8) issued by one of the two issuer warps; N2 = MMA columns; FIRST = first chunk
// of a bin (accumulators start from zero):
//   PROD 0  hi*hi, one MMA per k-step, k-step s into tile s
//   PROD 1  lo*hi and hi*lo, two MMAs per k-step into tile 2 (same issuing thread: program order)
template <uint32_t N2, int PROD, bool FIRST>
__device__ __forceinline__ void issue_chunk(uint32_t tmem, uint32_t tA, uint32_t blo) {
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 @17, M >> 4 @24
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N2 >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
#pragma unroll
  for (int k8 = 0; k8 < kTcChunk / 8; k8++) {
    const uint32_t a_hi = tA + k8 * 8, a_lo = a_hi + kTcChunk;
    const uint32_t b_hi = blo + k8 * ((2 * N2 * 16) >> 4), b_lo = b_hi + (kTcBHalf >> 4);
    if (PROD == 0) {
      if (FIRST) mma_ts<false>(tmem + k8 * kTcN2max, a_hi, b_hi, idesc);
      else mma_ts<true>(tmem + k8 * kTcN2max, a_hi, b_hi, idesc);
    } else {
      if (FIRST && k8 == 0) mma_ts<false>(tmem + 2 * kTcN2max, a_lo, b_hi, idesc);
      else mma_ts<true>(tmem + 2 * kTcN2max, a_lo, b_hi, idesc);
      mma_ts<true>(tmem + 2 * kTcN2max, a_hi, b_lo, idesc);
    }
  }
}

}  // namespace tc

template <int NLOG>
__global__ void __launch_bounds__(kTcThreads, 1) k_mimo_tc(MimoTcArgs a) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  using namespace tc;
  constexpr uint32_t N = 1u << NLOG;  // block-steps per CTA
  constexpr uint32_t N2 = 2 * N;      // columns of the accumulator tiles: (Xr | Xi)
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t kb = blockIdx.x * kTcBins, og = blockIdx.y, t0 = blockIdx.z * N;
  const uint32_t smem0 = smem_u32(tc_smem);
  float* tile = reinterpret_cast<float*>(tc_smem + kTcOffTile);
  const uint32_t raw0 = smem0 + kTcOffRaw;
  // mbarriers: full[NST] (producer arrivals), empty[NST] (tcgen05.commit), acc_full (commit), acc_empty (epilogue
  // threads), raw_full[RD] (loader + TMA bytes), raw_empty[RD] (producers)
  const uint32_t bar_full = smem0 + kTcOffBar;
  const uint32_t bar_empty = bar_full + 8 * kTcStages;
  const uint32_t bar_accf = bar_empty + 8 * kTcStages;
  const uint32_t bar_acce = bar_accf + 8;
  const uint32_t bar_rawf = bar_acce + 8;
  const uint32_t bar_rawe = bar_rawf + 8 * kTcRawStagesMax;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_smem + kTcOffBar + 8 * (2 * kTcStages + 2 + 2 * kTcRawStagesMax));
  const uint32_t RD = a.raw_stages, RSB = a.raw_stage_bytes;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTcStages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8 * s), "n"(kTcProducers));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_empty + 8 * s), "n"(kTcMmaWarps));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_accf), "n"(kTcMmaWarps));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_acce), "n"(kTcEpiThreads));
#pragma unroll
    for (int d = 0; d < kTcRawStagesMax; d++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_rawf + 8 * d));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_rawe + 8 * d), "n"(kTcProducers));
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
  const uint32_t P2log = a.P2log, P2 = 1u << P2log, P2m = P2 - 1;
  // raw B segment of one chunk: ninp input rows of seg complex each (rounded to 16 bytes), once for hi, once for lo
  const uint32_t ninp = P2log < 4 ? (16u >> P2log) : 1u;
  const uint32_t seg = P2log < 4 ? ((N + P2) & ~1u) : N + 16;
  const uint32_t bhalf = ninp * seg * 8;  // bytes of the hi (or lo) part of a chunk's FDL runs
  bool ok = true;
  const bool tr = a.trace != nullptr;
  unsigned long long* trow = tr ? a.trace + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 : nullptr;
  unsigned long long tw0 = 0, tw1 = 0;  // cycles this role waited on its two barriers
  const long long tstart = tr ? clock64() : 0;

  if (warp >= kTcWarpLoad) {
    // ================= loader: one lane issues the bulk copies (TMA) of the raw operands =================
    {
      // One bulk copy per chunk: the pre-staged FDL runs (k_mimo_pack_x), 2 * bhalf contiguous bytes.  The packed H does not
      // pass through shared memory at all (the producers read it from L2 into registers and prefetch it themselves).
      // This loop is a serial chain on one warp and paces the whole ring: running pointers only, no multiplies.
      const uint32_t bbytes = 2 * bhalf, total = kTcBins * nchunk;
      const char* xsrc = reinterpret_cast<const char*>(a.xb + (uint64_t)kb * a.xbin) + (uint64_t)blockIdx.z * nchunk * bbytes;
      const uint64_t bin_skip = a.xbin * 8 - (uint64_t)nchunk * bbytes;  // from the last chunk of a bin to the first of the next
      uint32_t d = 0, ph = 0, c = 0;
      uint32_t dst = raw0, bar = bar_rawf, bare = bar_rawe;
      const bool leader = elect_one();
      for (uint32_t it = 0; it < total; it++) {
        if (it >= RD) {
          if (ok && !mbar_wait_t(bare, ph ^ 1, tr, tw0)) ok = false;
        }
        if (leader) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bbytes) : "memory");
          bulk_g2s(dst, xsrc, bbytes, bar);
        }
        xsrc += bbytes;
        if (++c == nchunk) {
          c = 0;
          xsrc += bin_skip;
        }
        dst += RSB, bar += 8, bare += 8;
        if (++d == RD) {
          d = 0, ph ^= 1;
          dst = raw0, bar = bar_rawf, bare = bar_rawe;
        }
      }
      if (!ok && lane == 0) report_timeout(a.status, 3);
      if (tr && lane == 0) {
        trow[0] = (unsigned long long)(clock64() - tstart);  // loader: total, waiting for a free raw slot
        trow[1] = tw0;

      }
    }
  } else if (warp >= kTcWarpMma && warp < kTcWarpMma + kTcMmaWarps) {
    // ================= MMA issuer: the whole warp runs the loop, one elected lane issues =================
    // accumulator tiles 0, 1: hi*hi products of the chunk's two k-steps (warp 20); tile 2: the small lo*hi / hi*lo terms (warp 21)
    const uint32_t blo0 = kDescLo<N2> + (smem0 >> 4);  // low descriptor word of stage 0's B_hi tile
    uint32_t s = 0, ph = 0;
    for (uint32_t j = 0; j < (uint32_t)kTcBins; j++) {
      if (j >= 1 && ok && !mbar_wait_t(bar_acce, (j - 1) & 1, tr, tw1)) ok = false;  // the previous bin has been read out
      for (uint32_t c = 0; c < nchunk; c++) {
        if (ok && !mbar_wait_t(bar_full + 8 * s, ph, tr, tw0)) ok = false;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t tA = tmem + kTcAccCols + s * kTcAStageCols;
          const uint32_t blo = blo0 + s * (kTcStageBytes >> 4);
          if (warp == kTcWarpMma) {
            if (c == 0) issue_chunk<N2, 0, true>(tmem, tA, blo);
            else issue_chunk<N2, 0, false>(tmem, tA, blo);
          } else {
            if (c == 0) issue_chunk<N2, 1, true>(tmem, tA, blo);
            else issue_chunk<N2, 1, false>(tmem, tA, blo);
          }
          commit(bar_empty + 8 * s);
          if (c + 1 == nchunk) commit(bar_accf);
        }
        __syncwarp();
        if (++s == kTcStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
    if (!ok && lane == 0) report_timeout(a.status, 2);
    if (tr && lane == 0 && warp == kTcWarpMma) {
      trow[2] = (unsigned long long)(clock64() - tstart);  // issuer: total, waiting for operands, waiting for the read-out
      trow[3] = tw0;
      trow[4] = tw1;
    }
  } else if (warp >= kTcWarpEpi && warp < kTcWarpEpi + 4) {
    // ================= epilogue (4 warps, one per TMEM lane quarter) =================
    // Read out the accumulators of bin j and combine the quadrants.  Lane l < 16 of warp q holds the Hr row of output
    // o = 16 q + l, lane l + 16 the Hi row of the same output; per block-step t the first has (HrXr, HrXi) in columns t and
    // N + t, the second (HiXr, HiXi).  One shuffle of the Xi column gives re = HrXr - HiXi to the first lane and
    // im = HrXi + HiXr to the second.  Packed bin 0 = (DC, Nyquist) is two real products: re = HrXr, im = HiXi.
    // Results of two bins are collected in shared memory ([t][o][bin pair] float2) and stored 16 bytes per (t, output):
    // the four bins of a CTA fill one 32-byte sector of ypart[t][o][.] with two stores instead of four.
    const uint32_t q = warp & 3, et = tid - kTcWarpEpi * 32;
    const uint32_t part = lane >> 4, o = 16 * q + (lane & 15);
    for (uint32_t j = 0; j < (uint32_t)kTcBins; j++) {
      if (ok && !mbar_wait_t(bar_accf, j & 1, tr, tw0)) ok = false;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool bin0 = (kb + j) == 0;
      float* const outp = tile + (j & 1) * 2 + part;  // float index of (t = 0, o = 0) for this lane's component
#pragma unroll 1
      for (uint32_t cg = 0; cg < (N >> 3); cg++) {
        // columns cg*8 .. +7 (x Xr) and N + cg*8 .. +7 (x Xi) of this row in the three tiles: six loads in flight, one wait
        uint32_t r[2][kTcAccTiles][8];
#pragma unroll
        for (int xi = 0; xi < 2; xi++)
#pragma unroll
          for (uint32_t z = 0; z < kTcAccTiles; z++) {
            const uint32_t taddr = tmem + ((32 * q) << 16) + z * kTcN2max + cg * 8 + (xi ? N : 0u);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[xi][z][0]), "=r"(r[xi][z][1]), "=r"(r[xi][z][2]), "=r"(r[xi][z][3]), "=r"(r[xi][z][4]),
                           "=r"(r[xi][z][5]), "=r"(r[xi][z][6]), "=r"(r[xi][z][7])
                         : "r"(taddr));
          }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 8; u++) {
          // (t0 + t1) + t2 in fp32 round-to-nearest
          const float vr = (__uint_as_float(r[0][0][u]) + __uint_as_float(r[0][1][u])) + __uint_as_float(r[0][2][u]);
          const float vi = (__uint_as_float(r[1][0][u]) + __uint_as_float(r[1][1][u])) + __uint_as_float(r[1][2][u]);
          const float pvi = __shfl_xor_sync(0xffffffffu, vi, 16);
          // lane < 16 (Hr row): re = HrXr - HiXi;  lane >= 16 (Hi row): im = HrXi + HiXr
          float res = part ? pvi + vr : vr - pvi;
          if (bin0) res = part ? vi : vr;
          outp[((cg * 8 + u) * 64 + o) * 4] = res;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_acce);  // this thread's reads of the accumulators are complete: the next bin may start
      if (j & 1) {
        asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
        for (uint32_t idx = et; idx < N * 64; idx += kTcEpiThreads) {
          const uint32_t oo = idx & 63, t = idx >> 6;
          const uint32_t og_o = og * 64 + oo, tt = t0 + t;
          if (tt < a.T && og_o < a.n_out)
            *reinterpret_cast<float4*>(&a.ypart[((uint64_t)tt * a.slot_stride + og_o) * a.B + kb + j - 1]) =
                *reinterpret_cast<const float4*>(tile + (size_t)idx * 4);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");  // tile free for the next pair of bins
      }
    }
    if (!ok && lane == 0) report_timeout(a.status, 4);
    if (tr && et == 0) {
      trow[5] = (unsigned long long)(clock64() - tstart);  // epilogue: total, waiting for finished accumulators
      trow[6] = tw0;
    }
  } else if (warp < kTcGroups * kTcProducers / 32) {
    // ================= producers (4 groups of 4 warps): raw operands (smem) -> A tile (TMEM), B tile (smem) =====
    // Group g converts chunks it = g (mod 4) into operand stage g: four chunks are in flight at different points of
    // the chain (raw wait -> split -> TMEM / smem stores -> store wait + proxy fence -> hand-off).
    // A lives in tensor memory (row = lane, K along columns): warp w owns TMEM lanes 32 (w & 3) .. + 31 = rows m of
    // the planar matrix (m < 64: re of H for output m, m >= 64: im of H for output m - 64), the 16 K columns of the
    // chunk in two halves of 8.
    const uint32_t grp = warp >> 2, gtid = tid & (kTcProducers - 1);
    const uint32_t q = warp & 3;
    // row m = 32 q + lane of the planar matrix: lanes 0..15 = re of H for outputs 16 q .. + 15, lanes 16..31 = im of the same
    // outputs, so that the epilogue finds the two rows of an output in one warp (lane ^ 16)
    const uint32_t ao = 16 * q + (lane & 15);
    // packed H of this CTA's bins: chunk `it` (over all its bins) at hsrc + it * 8192; this row reads [re | im][g][ao]
    const char* hsrc = reinterpret_cast<const char*>(a.hpack + ((uint64_t)(og * a.B + kb) * a.G) * 64) + (lane >> 4) * 4096 + ao * 16;
    const uint32_t dstA = ((32 * q) << 16) + kTcAccCols + grp * kTcAStageCols;  // this group's stage (+ 16 for lo)
    // B (pre-split by k_mimo_pack_x): item e = gtid + 128 r -> (K group kg of four complex j, column t): four complex of
    // the raw runs -> the 16-byte row (kg, t) of their real parts and the row (kg, N + t) of their imaginary parts
    constexpr int NBR = (4 * N + kTcProducers - 1) / kTcProducers;  // 2, 1, 1
    uint32_t offB[NBR], srcB[NBR];
#pragma unroll
    for (int r = 0; r < NBR; r++) {
      const uint32_t e = gtid + kTcProducers * r;
      const uint32_t kg = (e >> NLOG) & 3, t = e & (N - 1), jl = kg << 2;
      offB[r] = kg * (N2 * 16) + t * 16;
      // element (local K index jl, column t) of the raw segment: row jl / P2, position t + p'
      srcB[r] = 8 * ((P2log < 4) ? (jl >> P2log) * seg + (jl & P2m) + t : jl + t);
    }
    // byte distance of the complex elements jl + 1, + 2, + 3 from jl (jl a multiple of four)
    const uint32_t dj1 = P2log == 0 ? 8 * seg : 8, dj2 = P2log == 0 ? 16 * seg : (P2log == 1 ? 8 * seg : 16),
                   dj3 = P2log == 0 ? 24 * seg : (P2log == 1 ? 8 * seg + 8 : 24);
    static_assert(kTcStages == kTcGroups, "one operand stage per producer group");
#ifdef BBX_TC_FINE_TRACE
    unsigned long long fa = 0, fb = 0, fc = 0, fd = 0;  // A part, B part, tcgen05.wait::st, fences + arrive
#endif
    const uint32_t total = kTcBins * nchunk;
    const uint32_t s = grp, sB = smem0 + s * kTcStageBytes;
    uint32_t ph = 0, d = grp % RD, phd = 0;
    // H of the group's first chunk; afterwards the loads of chunk it + 4 are issued as soon as the registers of chunk it
    // are free (after its TMEM stores) and land while the B tile is written
    float4 qa[4];
#pragma unroll
    for (int g = 0; g < 4; g++) qa[g] = ld_stream4(reinterpret_cast<const float4*>(hsrc + (uint64_t)grp * 8192 + g * 1024));
    // L2 prefetch of the 16 lines this warp reads per chunk ([re | im][g][256 bytes of outputs 16 q ..]): lane l < 16 takes
    // line l of the chunk kTcPrefetch ahead, one instruction per warp and chunk
    const char* const hpf = reinterpret_cast<const char*>(a.hpack + ((uint64_t)(og * a.B + kb) * a.G) * 64) + ((lane >> 3) & 1) * 4096 +
                            ((lane >> 1) & 3) * 1024 + q * 256 + (lane & 1) * 128;
    if (lane < 16) {
#pragma unroll
      for (uint32_t k = kTcGroups; k < (uint32_t)kTcPrefetch; k += kTcGroups)
        if (grp + k < kTcBins * nchunk) asm volatile("prefetch.global.L2 [%0];" ::"l"(hpf + (uint64_t)(grp + k) * 8192));
    }
    for (uint32_t it = grp; it < total; it += kTcGroups) {
      const uint32_t raw = raw0 + d * RSB;
      // ---- raw operands of this chunk have landed (TMA) ----
      if (ok && !mbar_wait_t(bar_rawf + 8 * d, phd, tr, tw0)) ok = false;
#ifdef BBX_TC_FINE_TRACE
      const long long f0 = clock64();
      const unsigned long long w1_before = tw1;
#endif
      // Everything that does not touch the operand stage comes first, so that the time between "stage free" and "stage
      // full" -- the part of a chunk that the four stages cannot hide -- is only the stores.
      // ---- A: the real (rows < 64) or imaginary (rows >= 64) parts of sixteen complex of output ao -> 16 K columns of row m
      float hi[16], lo[16];
#pragma unroll
      for (int g = 0; g < 4; g++) {
        split(qa[g].x, hi[4 * g], lo[4 * g]);
        split(qa[g].y, hi[4 * g + 1], lo[4 * g + 1]);
        split(qa[g].z, hi[4 * g + 2], lo[4 * g + 2]);
        split(qa[g].w, hi[4 * g + 3], lo[4 * g + 3]);
      }

      // ---- B, hi part of the first item: four complex of column t (the register budget of a 736-thread CTA ends here) ----
      float2 bh[4];
      bh[0] = ld_shared2(raw + srcB[0]), bh[1] = ld_shared2(raw + srcB[0] + dj1), bh[2] = ld_shared2(raw + srcB[0] + dj2),
      bh[3] = ld_shared2(raw + srcB[0] + dj3);
      if (it >= (uint32_t)kTcStages) {
        // ---- the MMAs that read this stage kTcStages chunks ago have completed ----
        if (ok && !mbar_wait_t(bar_empty + 8 * s, ph ^ 1, tr, tw1)) ok = false;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
#pragma unroll
      for (int half = 0; half < 2; half++) {
        float h8[8], l8[8];
#pragma unroll
        for (int u = 0; u < 8; u++) h8[u] = hi[8 * half + u], l8[u] = lo[8 * half + u];
        st_tmem8(tmem + dstA + 8 * half, h8);
        st_tmem8(tmem + dstA + 8 * half + kTcChunk, l8);
      }
      if (lane < 16 && it + kTcPrefetch < total)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(hpf + (uint64_t)(it + kTcPrefetch) * 8192));
      // the registers of this chunk's A values are free: H of the group's next chunk (lands while the B tile is written)
      if (it + kTcGroups < total) {
#pragma unroll
        for (int g = 0; g < 4; g++)
          qa[g] = ld_stream4(reinterpret_cast<const float4*>(hsrc + (uint64_t)(it + kTcGroups) * 8192 + g * 1024));
      }
#ifdef BBX_TC_FINE_TRACE
      const long long f1 = clock64();
#endif
      // ---- B: two 16-byte rows of the K-major tile per item (re | im), hi then lo ----
      const bool b_all = 4 * N >= kTcProducers * NBR;  // N = 16: half of the threads have an item
#pragma unroll
      for (int r = 0; r < NBR; r++) {
        if (b_all || gtid + kTcProducers * r < 4 * N) {
          const uint32_t src = raw + srcB[r], dst = sB + offB[r];
          if (r > 0) bh[0] = ld_shared2(src), bh[1] = ld_shared2(src + dj1), bh[2] = ld_shared2(src + dj2), bh[3] = ld_shared2(src + dj3);
          st_shared4(dst, bh[0].x, bh[1].x, bh[2].x, bh[3].x);
          st_shared4(dst + N * 16, bh[0].y, bh[1].y, bh[2].y, bh[3].y);
        }
      }
#pragma unroll
      for (int r = 0; r < NBR; r++) {
        if (b_all || gtid + kTcProducers * r < 4 * N) {
          const uint32_t src = raw + srcB[r] + bhalf, dst = sB + kTcBHalf + offB[r];
          const float2 c0 = ld_shared2(src), c1 = ld_shared2(src + dj1), c2 = ld_shared2(src + dj2), c3 = ld_shared2(src + dj3);
          st_shared4(dst, c0.x, c1.x, c2.x, c3.x);
          st_shared4(dst + N * 16, c0.y, c1.y, c2.y, c3.y);
        }
      }
      mbar_arrive(bar_rawe + 8 * d);  // raw slot free for the loader
      d += kTcGroups;
      while (d >= RD) {
        d -= RD;
        phd ^= 1;
      }
#ifdef BBX_TC_FINE_TRACE
      const long long f2 = clock64();
#endif
      // TMEM stores complete, smem writes visible to the tensor core (async proxy); hand the stage to the issuer
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#ifdef BBX_TC_FINE_TRACE
      const long long f3 = clock64();
#endif
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_full + 8 * s);
      ph ^= 1;
#ifdef BBX_TC_FINE_TRACE
      const long long f4 = clock64();
      fa += (unsigned long long)(f1 - f0) - (tw1 - w1_before);
      fb += (unsigned long long)(f2 - f1);
      fc += (unsigned long long)(f3 - f2);
      fd += (unsigned long long)(f4 - f3);
#endif
    }
    if (!ok && gtid == 0) report_timeout(a.status, 1);
#ifdef BBX_TC_FINE_TRACE
    if (tr && gtid == 0 && grp == 0) {
      trow[7] = (unsigned long long)(clock64() - tstart);
      trow[8] = tw0;
      trow[9] = tw1;
      trow[10] = fa;
      trow[11] = fb;
      trow[12] = fc;
      trow[13] = fd;
    }
#else
    if (tr && gtid == 0) {
      trow[7 + 3 * grp] = (unsigned long long)(clock64() - tstart);  // producer group: total, raw wait, stage wait
      trow[8 + 3 * grp] = tw0;
      trow[9 + 3 * grp] = tw1;
    }
#endif
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTcTmemCols));
}

// Xq (layout above) from the FDL ring: one (column tile, chunk, input row) segment per blockIdx.y, 32 bins x 32
// positions per CTA through a shared-memory transpose (FDL rows are bin-contiguous, segments time-contiguous); the
// TF32 hi / lo split happens here so that the GEMM kernel only copies its B operand.
__global__ void __launch_bounds__(256) k_mimo_pack_x(const float2* __restrict__ fdl, float2* __restrict__ xq, uint32_t B,
                                                     uint32_t R, uint32_t head, uint32_t n_in, uint32_t P2log, uint32_t T,
                                                     uint32_t N, uint32_t nchunk, uint64_t xbin) {
  __shared__ float2 tile[32][33];
  const uint32_t P2 = 1u << P2log;
  const uint32_t ninp = P2log < 4 ? (16u >> P2log) : 1u;
  const uint32_t seg = P2log < 4 ? ((N + P2) & ~1u) : N + 16;
  // segment -> (column tile tt, chunk c, row il of the chunk) -> input i, first reversed partition p'0
  const uint32_t il = blockIdx.y % ninp, c = (blockIdx.y / ninp) % nchunk, tt = blockIdx.y / (ninp * nchunk);
  const uint32_t i = P2log < 4 ? c * ninp + il : (c >> (P2log - 4));
  const uint32_t pp0 = P2log < 4 ? 0u : ((c << 4) & (P2 - 1));
  const uint32_t k0 = blockIdx.x * 32, w0 = blockIdx.z * 32;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t wl = w0 + ty + 8 * r;
    float2 v = make_float2(0.f, 0.f);
    // block of this call (negative: history) met by column t = wl - p' at reversed partition p'
    const long long blk = (long long)tt * N + pp0 + wl - (long long)(P2 - 1);
    if (wl < seg && i < n_in && blk < (long long)T) {
      long long sl = ((long long)head + blk) % (long long)R;
      if (sl < 0) sl += R;
      v = fdl[((uint64_t)i * R + (uint64_t)sl) * B + k0 + tx];
    }
    tile[ty + 8 * r][tx] = v;
  }
  __syncthreads();
  const uint64_t chunk_elems = 2ull * ninp * seg;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t kl = ty + 8 * r, wl = w0 + tx;
    if (wl < seg) {
      const float2 v = tile[tx][kl];
      float2 h, l;
      tc::split(v.x, h.x, l.x);
      tc::split(v.y, h.y, l.y);
      const uint64_t at = (uint64_t)(k0 + kl) * xbin + ((uint64_t)tt * nchunk + c) * chunk_elems + (uint64_t)il * seg + wl;
      xq[at] = h;
      xq[at + (uint64_t)ninp * seg] = l;
    }
  }
}

// Hpack[og][k][chunk][part][g][o] = the real (part 0) / imaginary (part 1) parts of H_{o,i}[p][k] for the four complex K
// indices j = 16 chunk + 4 g .. + 3 (j = i*P2 + p', p = P2-1-p'); zero where the filter is null / shorter / beyond n_in.
// One CTA: 32 bins x one group of four j x 64 outputs (two passes of 32 outputs).  Runs when the filter matrix changes.
__global__ void __launch_bounds__(256) k_mimo_pack_h(const float2* const* __restrict__ ftab, const uint32_t* __restrict__ fparts,
                                                     float4* __restrict__ hpack, uint32_t B, uint32_t n_in, uint32_t n_out,
                                                     uint32_t P2log, uint32_t G) {
  __shared__ float2 tile[32][129];  // [k][o_local * 4 + jj]
  const uint32_t k0 = blockIdx.x * 32, g4 = blockIdx.y, og = blockIdx.z;  // g4 = group of four complex K indices
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t P2 = 1u << P2log;
  const uint32_t chunk = g4 >> 2, gl = g4 & 3;
  for (uint32_t oh = 0; oh < 2; oh++) {
    for (uint32_t pr = warp; pr < 128; pr += 8) {  // (o_local, jj) pairs, lanes over k
      const uint32_t o = 32 * oh + (pr >> 2), jj = pr & 3;
      const uint32_t j = 4 * g4 + jj, i = j >> P2log, p = P2 - 1 - (j & (P2 - 1));
      const uint32_t og_o = og * 64 + o;
      float2 v = make_float2(0.f, 0.f);
      if (og_o < n_out && i < n_in) {
        const float2* H = ftab[(uint64_t)og_o * n_in + i];
        if (H && p < fparts[(uint64_t)og_o * n_in + i]) v = H[(uint64_t)p * B + k0 + lane];
      }
      tile[lane][pr] = v;
    }
    __syncthreads();
    for (uint32_t kl = warp; kl < 32; kl += 8) {
      const float2 c0 = tile[kl][4 * lane], c1 = tile[kl][4 * lane + 1], c2 = tile[kl][4 * lane + 2], c3 = tile[kl][4 * lane + 3];
      // float4 index of (og, k, chunk, part, gl, o): G float4 groups of 64 per (og, k) = chunks x 2 parts x 4 groups
      const uint64_t at = ((((uint64_t)og * B + k0 + kl) * (G / 8) + chunk) * 8 + gl) * 64 + 32 * oh + lane;
      hpack[at] = make_float4(c0.x, c1.x, c2.x, c3.x);
      hpack[at + 4 * 64] = make_float4(c0.y, c1.y, c2.y, c3.y);
    }
    __syncthreads();
  }
}

}  // namespace bbx
