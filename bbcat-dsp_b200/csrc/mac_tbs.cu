// mac_tbs.cu -- k_fdl_mac_tbs: the time-batched FDL multiply-accumulate with operands streamed through shared memory
// by a TMA producer warp (the dominant kernel of any call with >= 8 blocks and long filters), and k_nyq_mac2, the
// Nyquist side sums that go with it.
//
// Same sums, same plan, same per-output FMA order as k_fdl_mac / k_fdl_mac_tb (kernels_mac.cuh): bit-identical results.
// What changed against k_fdl_mac_tb (round 1: 0.60 of the nominal FP32 rate, FMA pipe 70 % busy):
//   * per (channel, bin) the sums over a call's T block-steps are a length-P FIR along the block axis.  A thread owns one
//     bin and TT = 16 consecutive block-steps (16 float2 accumulators + a 16-row register window of the FDL that is
//     rotated by static indexing); per partition it needs ONE filter value and ONE new FDL value for 32 packed FMAs.
//   * round 1 fetched those two values per thread with cp.async (LDGSTS): two 8-byte copies + address arithmetic per
//     32 FFMA2, 21 non-FMA instructions per step, and the LDGSTS rate of the SM (8 cycles per warp instruction) was as
//     much a bound as the FMA pipe.  Here a CTA is 64 columns x 4 time tiles (256 consumer threads): the four tiles
//     share every filter row and every FDL row.  Rows travel into shared-memory rings in chunks of 4 KB, ONE 16-byte
//     cp.async per thread per chunk (256 threads = one chunk), completion counted on mbarriers
//     (cp.async.mbarrier.arrive): 0.25 LDGSTS per thread and step instead of 2.  Consumers issue two LDS.64 with
//     immediate offsets per step; L2 -> SM traffic per FMA drops 3.4x.  (A first version fed the rings with one
//     512-byte cp.async.bulk per row from a producer warp: 715 k bulk copies per launch, one per ~160 cycles and SM --
//     the TMA unit's rate for small copies was the bound and the kernel ran at 0.42 ms; ncu in profiles/.)
//   * the register window of a new (channel, partition range) segment is filled from the same shared-memory stream
//     (the segment's FDL rows simply start TT * NTILE - 1 rows early), and the copies run ahead across segment
//     boundaries, so the window fill and the pipeline fill of round 1 (~20 % of a CTA's time) overlap the previous
//     segment's arithmetic.
//   * persistent CTAs: one CTA pair per SM walks several consecutive row ranges of the plan, so there is one ramp-up
//     per launch instead of one per wave.
#include <stdlib.h>

#include <algorithm>

#include "async_copy.cuh"
#include "mac_common.cuh"
#include "mac_tbs.h"
#include "mac_tbw.cuh"

namespace bbx {

template <int NTILE, int NTHREADS, int GPS>
struct TbsCfg {
  static constexpr int TT = 16;                      // block-steps per thread
  static constexpr int COLS = NTHREADS / NTILE;      // bins per CTA
  static constexpr int ROWB = COLS * 8;              // bytes of one row tile
  static constexpr int CHB = NTHREADS * 16;          // bytes per chunk: every thread copies one 16-byte piece
  static constexpr int CH = CHB / ROWB;              // rows per chunk: 8 (NTILE = 4), 4 (NTILE = 2)
  static constexpr int GC = TT / CH;                 // chunks per group of 16 steps
  static constexpr int FILL = TT * NTILE;            // FDL rows a segment needs before its first step
  static constexpr int FILLC = FILL / CH;
  static constexpr int XCH = 16;                     // chunks of the FDL ring: the live window + one group + look-ahead
  static constexpr int HCH = 8;                      // chunks of the filter ring: one group + look-ahead
  static constexpr int THREADS = NTHREADS;
  static constexpr int CTAS_PER_SM = 512 / NTHREADS;  // 16 warps per SM: the register file allows no more at ~120 registers
  static constexpr int SMEM = (XCH + HCH) * CHB;
  static_assert(NTILE == 2 || NTILE == 4, "tiles per CTA");
  static constexpr int IC = GC * GPS;                // chunks per interval between two synchronisation points
  // during an interval the FDL ring holds the live window (FILLC - GC + 1 chunks) and the interval's IC chunks, and the next
  // interval's IC chunks are being copied in; the filter ring this interval's and the next one's chunks
  static_assert(FILLC - GC + 1 + 2 * IC <= XCH && 2 * IC <= HCH, "ring sizes");
};

// Chunk streams.  Every segment (a run of partitions p0 .. p0 + np - 1 of one filter against one input's FDL) is two
// row streams, each cut into chunks of CH rows starting at a fresh chunk (the last chunk of a segment may be short):
//   filter stream  q = 0 .. np - 1          row p0 + q of H
//   FDL stream     j = 0 .. FILL - 2 + np   row (base + FILL - 1 - j) mod R, base = slot of the tile group's first step - p0
// Tile i (block-steps 16 i .. 16 i + 15 of the group) meets, at step q, FDL stream row j = FILL - 1 - 16 i + q; its
// register window holds the 15 rows before that.  Chunks are issued in the order they are needed: the FILLC chunks of
// the fill, then filter chunk k / FDL chunk FILLC + k alternately; chunk c of a stream lives in ring slot c mod ring size.
//
// Synchronisation is one block barrier per group of 16 steps and nothing else.  At the top of a group every thread
// waits for its own copies (cp.async.wait_all), the barrier makes everybody's pieces visible and proves that everybody
// has finished the previous group, and then every thread issues its 16 bytes of the following chunks for as far as
// the rings have room (the chunks all tiles have finished with are known by construction, all warps being at the same
// step).  The chunks a group reads were therefore issued a whole group earlier -- 16 steps x 32 packed FMAs per warp to
// cover the L2 / HBM latency -- and the inner loop is LDS + FFMA2 only.  (Two mbarrier-based versions came first: a
// TMA producer warp with one 512-byte cp.async.bulk per row, 0.42 ms, and per-thread cp.async with
// cp.async.mbarrier.arrive, 0.31 ms: try_wait / arrive on SM 10.0 cost ~12 instructions each -- the cluster-window
// address is rebuilt from SR_CgaCtaId --, 16 to 20 barrier operations per group came to 700 instructions next to the 512
// FFMA2; ncu captures in profiles/.)
template <int NTILE, int NTHREADS, int GPS>
__global__ void __launch_bounds__(NTHREADS, 512 / NTHREADS)
k_fdl_mac_tbs(const MacSeg* __restrict__ segs, const uint32_t* __restrict__ cta_seg_begin, uint32_t n_plan_ctas,
              uint32_t plan_per_cta, const float2* __restrict__ fdl, float2* __restrict__ ypart, uint32_t B, uint32_t R,
              uint32_t head0, uint32_t t0, uint32_t nt, uint32_t ncoltiles, uint32_t slot_stride, int* __restrict__ status) {
  using C = TbsCfg<NTILE, NTHREADS, GPS>;
  constexpr int TT = C::TT, COLS = C::COLS, ROWB = C::ROWB, CH = C::CH, CHB = C::CHB, FILL = C::FILL, FILLC = C::FILLC;
  constexpr int XCH = C::XCH, HCH = C::HCH, GC = C::GC, IC = C::IC;
  extern __shared__ __align__(128) uint8_t tbs_smem[];
  const uint32_t xring = (uint32_t)__cvta_generic_to_shared(tbs_smem), hring = xring + XCH * CHB;
  (void)status;

  const uint32_t coltile = blockIdx.x % ncoltiles, tgroup = blockIdx.x / ncoltiles;
  const uint32_t tbase0 = tgroup * (TT * NTILE);     // first block-step of this tile group, relative to t0
  const uint32_t s0 = (head0 + t0 + tbase0) % R;     // its FDL slot
  const uint32_t col0 = coltile * COLS;
  const uint32_t pc0 = blockIdx.y * plan_per_cta, pc1 = min(pc0 + plan_per_cta, n_plan_ctas);
  if (pc0 >= pc1) return;
  const uint32_t sb = cta_seg_begin[pc0], se = cta_seg_begin[pc1];

  // ---- copy cursors (identical in every thread): each stream is walked on its own with running pointers ----
  constexpr uint32_t kPiecesPerRow = ROWB / 16;
  const uint32_t prow = threadIdx.x / kPiecesPerRow;  // my row of every chunk
  const uint32_t my16 = threadIdx.x * 16;             // offset of my 16 bytes inside a chunk's slot
  const uint32_t poff = (threadIdx.x % kPiecesPerRow) * 16;
  const uint32_t chunk_bytes_g = CH * B * (uint32_t)sizeof(float2);  // global bytes between my pieces of consecutive chunks
  const uint64_t ring_bytes_g = (uint64_t)R * B * sizeof(float2);
  uint32_t px_si = sb, ph_si = sb;  // segment each stream is in
  uint32_t p_xn = 0, p_hn = 0;      // chunks issued since the kernel started = global index of the next chunk
  int p_xrem = 0, p_hrem = 0;       // rows of the segment's stream not yet issued (from the first row of the next chunk)
  int p_xr = 0;                     // FDL ring row of my piece of the next FDL chunk
  const char* p_xptr = nullptr;     // ... and its address
  const char* p_hptr = nullptr;     // address of my piece of the next filter chunk
  auto open_x = [&]() {
    const MacSeg sg = segs[px_si];
    p_xrem = (int)(FILL - 1 + sg.np);
    uint32_t base = s0 + R - (sg.p0 % R);
    if (base >= R) base -= R;
    int r = (int)(base + FILL - 1) - (int)prow;  // >= 0; short rings (R < FILL: few blocks per call on 64 columns) wrap twice
    while (r >= (int)R) r -= (int)R;
    p_xr = r;
    p_xptr = reinterpret_cast<const char*>(fdl + ((uint64_t)sg.fdl_ch * R + (uint32_t)r) * B + col0) + poff;
  };
  auto open_h = [&]() {
    const MacSeg sg = segs[ph_si];
    p_hrem = (int)sg.np;
    p_hptr = reinterpret_cast<const char*>(reinterpret_cast<const float2*>(sg.H) + (uint64_t)(sg.p0 + prow) * B + col0) + poff;
  };
  if (sb < se) {
    open_x();
    open_h();
  }
  // One synchronisation point: the chunks below (need_x, need_h) are about to be read; the chunks below (done_x, done_h)
  // have been read by every tile.  Issues this thread's pieces of the following chunks: up to kLook intervals beyond what
  // this interval reads, as far as the rings have room -- across segment boundaries (the next segment's fill comes next
  // in the FDL stream).
  constexpr uint32_t kLook = 3;
  auto sync_point = [&](uint32_t need_x, uint32_t need_h, uint32_t done_x, uint32_t done_h) {
    ac::cp_async_wait_all();
    __syncthreads();
    const bool late = p_xn < need_x || p_hn < need_h;  // the start of the kernel; a fill the FDL ring had no room for earlier
    const uint32_t tx = min(need_x + kLook * IC, done_x + XCH), th = min(need_h + kLook * IC, done_h + HCH);
    while (p_xn < tx && px_si < se) {
      if (p_xrem > (int)prow) ac::cp_async16(xring + (p_xn & (XCH - 1)) * CHB + my16, p_xptr);
      p_xn++;
      p_xrem -= CH;
      p_xr -= CH;
      p_xptr -= chunk_bytes_g;
      if (p_xr < 0) {  // the ring wraps
        p_xr += (int)R;
        p_xptr += ring_bytes_g;
      }
      if (p_xrem <= 0) {
        px_si++;
        if (px_si < se) open_x();
      }
    }
    while (p_hn < th && ph_si < se) {
      if (p_hrem > (int)prow) ac::cp_async16(hring + (p_hn & (HCH - 1)) * CHB + my16, p_hptr);
      p_hn++;
      p_hrem -= CH;
      p_hptr += chunk_bytes_g;
      if (p_hrem <= 0) {
        ph_si++;
        if (ph_si < se) open_h();
      }
    }
    if (late) {
      ac::cp_async_wait_all();
      __syncthreads();
    }
  };

  // ==================================== thread = (bin, tile of 16 block-steps) ========================================
  const uint32_t col = threadIdx.x % COLS, tile = threadIdx.x / COLS;
  const uint32_t xbase = xring + col * 8, hbase = hring + col * 8;
  // FDL chunk (relative to the segment's first) that holds this tile's row of step 0 -- its last row
  const uint32_t xfirst = (FILL - TT * tile) / CH - 1;
  float2 acc[TT], W[TT];
#pragma unroll
  for (int i = 0; i < TT; i++) acc[i] = make_float2(0.f, 0.f);
  uint32_t xc = 0, hc = 0;  // first chunk of the current segment's streams
  for (uint32_t si = sb; si < se; si++) {
    const MacSeg sg = segs[si];
    const uint32_t np = sg.np;
    const uint32_t nxc = (FILL - 1 + np + CH - 1) / CH, nhc = (np + CH - 1) / CH;
    if (sg.flags & 1u) {
#pragma unroll
      for (int i = 0; i < TT; i++) acc[i] = make_float2(0.f, 0.f);
    }
    uint32_t xcur = 0, hcur = 0;
    float2 hn = make_float2(0.f, 0.f), xn = hn;  // operands of the next step, loaded under the current step's FMAs
    // q + u = step of the segment; u = (q + u) mod 16 because groups start at multiples of 16, so chunk boundaries fall
    // on fixed u.  Filter chunk of the step: hc + (q + u) / CH.  My FDL chunk: xc + xfirst + (q + u + CH - 1) / CH.
    auto fetch = [&](const int u, const uint32_t q) {
      if (u % CH == 0) hcur = hbase + ((hc + q / CH + u / CH) & (HCH - 1)) * CHB;
      if (u % CH == 1 || u == 0) xcur = xbase + ((xc + xfirst + (q + u + CH - 1) / CH) & (XCH - 1)) * CHB;
      hn = ac::lds2(hcur + (u % CH) * ROWB);
      xn = ac::lds2(xcur + ((u + CH - 1) % CH) * ROWB);
    };
    auto step = [&](const int u, const uint32_t q, const bool more) {
      const float2 h = hn;
      W[(TT - u) % TT] = xn;
      if (more && u + 1 < TT) fetch(u + 1, q);  // the loads of step u + 1 fly under the 32 FFMA2 of step u
#pragma unroll
      for (int i = 0; i < TT; i++) cmac_x2(acc[i], h, W[(i - u + TT) % TT]);
    };
    for (uint32_t q0 = 0; q0 < np; q0 += TT * GPS) {
      // the interval q0 .. q0 + 16 GPS - 1 reads filter chunks below q0 / CH + IC and FDL chunks below FILLC + q0 / CH + IC;
      // every tile is through with the filter chunks below q0 / CH and (after the first group) the FDL chunks below
      // GC - 1 + q0 / CH
      sync_point(xc + min(nxc, FILLC + q0 / CH + IC), hc + min(nhc, q0 / CH + IC), xc + (q0 ? GC - 1 + q0 / CH : 0u), hc + q0 / CH);
#pragma unroll
      for (int g = 0; g < GPS; g++) {
        const uint32_t q = q0 + g * TT;
        if (q >= np) break;
        fetch(0, q);
        if (q == 0) {
          // W[e] = row base_i + e = stream row FILL - 1 - 16 i - e (e = 1 .. 15); W[0] is overwritten by step 0
          const uint32_t j0 = FILL - TT * (tile + 1);  // stream row of e = 15
#pragma unroll
          for (int e = 1; e < TT; e++) {
            const uint32_t j = j0 + (TT - 1 - e);
            W[e] = ac::lds2(xbase + ((xc + j / CH) & (XCH - 1)) * CHB + (j % CH) * ROWB);
          }
          W[0] = make_float2(0.f, 0.f);
        }
        if (q + TT <= np) {
#pragma unroll
          for (int u = 0; u < TT; u++) step(u, q, true);
        } else {
#pragma unroll
          for (int u = 0; u < TT; u++)
            if (q + u < np) step(u, q, q + u + 1 < np);
        }
      }
    }
    xc += nxc;
    hc += nhc;
    if (sg.flags & 2u) {
      const uint32_t tb = tbase0 + tile * TT;
#pragma unroll
      for (int i = 0; i < TT; i++)
        if (tb + i < nt) ypart[((uint64_t)(tb + i) * slot_stride + sg.slot) * B + col0 + col] = acc[i];
    }
  }
  ac::cp_async_wait_all();
}

// Nyquist sums next to the time-batched MAC: N[t][run] = sum over the run's rows of Nqh[p] * Nqx[s_t - p] (the
// imaginary parts of column 0), p ascending, one fma per row -- the sequence the streaming kernel runs inline, hence the
// same bits.  Column 0 of consecutive rows is 8 B bytes apart, so the values of a piece of a segment are first gathered
// into shared memory by the whole CTA (independent loads, all in flight together) and the serial chains then run from
// there; thread = block-step.  (Round 1's k_nyq_mac chased those strided loads inside the chain: 28 us.)
static constexpr int kNyqPiece = 256;
__global__ void __launch_bounds__(64)
k_nyq_mac2(const MacSeg* __restrict__ segs, const uint32_t* __restrict__ cta_seg_begin, const float2* __restrict__ fdl,
           float* __restrict__ nyq_part, uint32_t B, uint32_t R, uint32_t head0, uint32_t t0, uint32_t nt, uint32_t slot_stride) {
  __shared__ float hs[kNyqPiece];
  __shared__ float xs[kNyqPiece + 64];
  const uint32_t tl = threadIdx.x, tq = blockIdx.y * 64 + tl;
  const uint32_t sbase = (head0 + t0 + blockIdx.y * 64) % R;  // FDL slot of this tile's first block-step
  const uint32_t sb = cta_seg_begin[blockIdx.x], se = cta_seg_begin[blockIdx.x + 1];
  float acc = 0.f;
  for (uint32_t si = sb; si < se; si++) {
    const MacSeg sg = segs[si];
    if (sg.flags & 1u) acc = 0.f;
    const float2* hcol = reinterpret_cast<const float2*>(sg.H) + (uint64_t)sg.p0 * B;
    const float2* xb = fdl + (uint64_t)sg.fdl_ch * R * B;
    uint32_t base = sbase + R - (sg.p0 % R);
    if (base >= R) base -= R;
    for (uint32_t pp = 0; pp < sg.np; pp += kNyqPiece) {
      const uint32_t len = min((uint32_t)kNyqPiece, sg.np - pp);
      __syncthreads();  // the previous piece has been consumed
      for (uint32_t k = tl; k < len; k += 64) hs[k] = __ldg(&hcol[(uint64_t)(pp + k) * B]).y;
      // xs[k] = row (base - pp - (len - 1) + k) mod R, k = 0 .. len + 62: step p of lane t reads xs[t + len - 1 - (p - pp)]
      for (uint32_t k = tl; k < len + 63; k += 64) {
        int r = (int)base - (int)pp - (int)(len - 1) + (int)k;
        r %= (int)R;
        if (r < 0) r += (int)R;
        xs[k] = __ldg(&xb[(uint64_t)r * B]).y;
      }
      __syncthreads();
      const float* xp = xs + tl + len - 1;
#pragma unroll 8
      for (uint32_t k = 0; k < len; k++) acc = fmaf(hs[k], xp[-(int)k], acc);
    }
    if ((sg.flags & 2u) && tq < nt) nyq_part[(uint64_t)tq * slot_stride + sg.slot] = acc;
  }
}

template <int NTILE, int NTHREADS, int GPS>
static cudaError_t launch_tbs_t(const MacTbsArgs& a, cudaStream_t st) {
  using C = TbsCfg<NTILE, NTHREADS, GPS>;
  static uint32_t attr_set = 0;  // per device: function attributes belong to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set & (1u << (dev & 31)))) {
    cudaError_t e = cudaFuncSetAttribute(k_fdl_mac_tbs<NTILE, NTHREADS, GPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set |= 1u << (dev & 31);
  }
  const uint32_t ncol = a.B / C::COLS, ngroups = ceil_div(a.nt, (uint32_t)(C::TT * NTILE));
  // persistent: one resident wave of CTAs, each walking consecutive row ranges of the plan
  uint32_t gy = std::max(1u, ((uint32_t)C::CTAS_PER_SM * kNumSMs) / (ncol * ngroups));
  gy = std::min(gy, a.n_plan_ctas);
  const uint32_t per = ceil_div(a.n_plan_ctas, gy);
  gy = ceil_div(a.n_plan_ctas, per);
  k_fdl_mac_tbs<NTILE, NTHREADS, GPS><<<dim3(ncol * ngroups, gy), C::THREADS, C::SMEM, st>>>(
      a.segs, a.cta_seg_begin, a.n_plan_ctas, per, a.fdl, a.ypart, a.B, a.R, a.head, a.t0, a.nt, ncol, a.slot_stride, a.status);
  return cudaGetLastError();
}

cudaError_t launch_nyq_mac2(const MacTbsArgs& a, cudaStream_t st) {
  // The side kernel asks for the shared-memory carve-out the MAC needs (all of it).  An SM changes its carve-out only when
  // it is empty: where a side CTA had been placed first under a smaller carve-out, the MAC's CTAs (102 KB each) had to wait
  // for it to leave -- the MAC launch then took the side kernel's 12 us longer.  With the same carve-out they join it.
  static uint32_t carve_set = 0;  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(carve_set & (1u << (dev & 31)))) {
    cudaFuncSetAttribute(k_nyq_mac2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carve_set |= 1u << (dev & 31);
  }
  k_nyq_mac2<<<dim3(a.n_plan_ctas, ceil_div(a.nt, 64u)), 64, 0, st>>>(a.segs, a.cta_seg_begin, a.fdl, a.nyq_part, a.B, a.R, a.head, a.t0,
                                                                      a.nt, a.slot_stride);
  return cudaGetLastError();
}

static cudaError_t launch_tbw(const MacTbsArgs& a, cudaStream_t st) {
  using C = TbwCfg;
  static uint32_t attr_set = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set & (1u << (dev & 31)))) {
    cudaError_t e = cudaFuncSetAttribute(k_fdl_mac_tbw, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set |= 1u << (dev & 31);
  }
  const uint32_t ncol = a.B / C::COLS, ngroups = ceil_div(a.nt, (uint32_t)(C::TT * C::NTILE));
  // persistent: 16 warps per SM in total, each walking consecutive row ranges of the plan
  uint32_t ranges = std::max(1u, (2u * C::WARPS * kNumSMs) / (ncol * ngroups));
  ranges = std::min(ranges, a.n_plan_ctas);
  const uint32_t per = ceil_div(a.n_plan_ctas, ranges);
  ranges = ceil_div(a.n_plan_ctas, per);
  const uint32_t nwarps = ranges * ncol * ngroups;
  k_fdl_mac_tbw<<<ceil_div(nwarps, (uint32_t)C::WARPS), 32 * C::WARPS, C::SMEM, st>>>(a.segs, a.cta_seg_begin, a.n_plan_ctas, per, ranges,
                                                                                   a.fdl, a.ypart, a.B, a.R, a.head, a.t0, a.nt, ncol,
                                                                                   ngroups, a.slot_stride);
  return cudaGetLastError();
}

cudaError_t launch_mac_tbs(const MacTbsArgs& a, cudaStream_t st, const char** kernel_name) {
  // tiles per CTA: four (64 block-steps per CTA) unless the call is short; CTA size: BBX_TBS_THREADS for A/B runs
  static int threads = 0, force_tiles = 0, gps = 0;
  if (!threads) {
    const char* t = getenv("BBX_TBS_THREADS");
    threads = (t && atoi(t) == 256) ? 256 : 128;
    const char* n = getenv("BBX_TBS_NTILE");
    force_tiles = n ? atoi(n) : 0;
    const char* g = getenv("BBX_TBS_GPS");
    gps = (g && atoi(g) == 1) ? 1 : 2;
  }
  static int use_tbw = -1;
  if (use_tbw < 0) {
    const char* w = getenv("BBX_TBW");
    use_tbw = (w && atoi(w) == 0) ? 0 : 1;
  }
  if (use_tbw && a.nt > 32) {  // warp-private streams: four tiles of 16 block-steps per warp
    *kernel_name = "k_fdl_mac_tbw";
    return launch_tbw(a, st);
  }
  bool four = a.nt > 32;
  if (force_tiles == 2) four = false;
  if (force_tiles == 4) four = true;
  if (threads == 256 && a.B < 128) four = true;  // the column tile (256 / tiles) must fit the block size
#define BBX_TBS_CASE(NT_, TH_, G_)                       \
  do {                                                    \
    *kernel_name = "k_fdl_mac_tbs<" #NT_ "," #TH_ "," #G_ ">"; \
    return launch_tbs_t<NT_, TH_, G_>(a, st);              \
  } while (0)
  if (threads == 256) {
    if (four) {
      if (gps == 1) BBX_TBS_CASE(4, 256, 1);
      BBX_TBS_CASE(4, 256, 2);
    }
    BBX_TBS_CASE(2, 256, 1);
  }
  if (four) {
    if (gps == 1) BBX_TBS_CASE(4, 128, 1);
    BBX_TBS_CASE(4, 128, 2);
  }
  BBX_TBS_CASE(2, 128, 1);
#undef BBX_TBS_CASE
}

}  // namespace bbx
