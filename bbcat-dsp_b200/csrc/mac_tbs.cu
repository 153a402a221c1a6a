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
//     share every filter row and every FDL row, which ONE producer warp streams into shared-memory rings with
//     cp.async.bulk (one 512-byte bulk copy per row tile, completion on mbarriers).  Consumers issue two LDS.64 with
//     immediate offsets per step and nothing else; L2 -> SM traffic per FMA drops 3.4x.
//   * the register window of a new (channel, partition range) segment is filled from the same shared-memory stream
//     (the producer simply starts the segment's FDL rows TT * NTILE - 1 rows early), and the producer runs ahead across
//     segment boundaries, so the window fill and the pipeline fill of round 1 (~20 % of a CTA's time) overlap the
//     previous segment's arithmetic.
//   * persistent CTAs: one CTA pair per SM walks several consecutive row ranges of the plan, so there is one ramp-up
//     per launch instead of one per wave.
#include <algorithm>

#include "async_copy.cuh"
#include "mac_common.cuh"
#include "mac_tbs.h"

namespace bbx {

template <int NTILE>
struct TbsCfg {
  static constexpr int TT = 16;                      // block-steps per thread
  static constexpr int COLS = 256 / NTILE;           // bins per CTA
  static constexpr int ROWB = COLS * 8;              // bytes of one row tile
  static constexpr int CHB = 4096;                   // bytes per chunk = one mbarrier phase
  static constexpr int CH = CHB / ROWB;              // rows per chunk: 8 (NTILE = 4), 4 (NTILE = 2)
  static constexpr int GC = TT / CH;                 // chunks per group of 16 steps
  static constexpr int FILL = TT * NTILE;            // FDL rows a segment needs before its first step
  static constexpr int FILLC = FILL / CH;
  static constexpr int XCH = 16;                     // chunks of the FDL ring: the live window + one group + look-ahead
  static constexpr int HCH = 8;                      // chunks of the filter ring: one group + look-ahead
  static constexpr int LA = (NTILE == 4) ? 4 : 3;    // chunks the producer tries to stay ahead of the current group
  static constexpr int NCONS = 8;                    // consumer warps (the last one doubles as the producer)
  static constexpr int THREADS = 32 * NCONS;
  static constexpr int BAR_BYTES = 8 * 2 * (XCH + HCH) + 8;  // + the "a wait timed out" word
  static constexpr int SMEM = (XCH + HCH) * CHB + BAR_BYTES;
  static_assert(NTILE == 2 || NTILE == 4, "tiles per CTA");
  static_assert(FILLC - (TT / CH) + 1 + GC + LA <= XCH, "FDL ring: live window + current group + look-ahead");
  static_assert(GC + LA <= HCH, "filter ring: current group + look-ahead");
};

// Chunk streams.  Every segment (a run of partitions p0 .. p0 + np - 1 of one filter against one input's FDL) is two
// row streams, each cut into chunks of CH rows starting at a fresh chunk (the last chunk of a segment may be short):
//   filter stream  q = 0 .. np - 1          row p0 + q of H
//   FDL stream     j = 0 .. FILL - 2 + np   row (base + FILL - 1 - j) mod R, base = slot of the tile group's first step - p0
// Tile i (block-steps 16 i .. 16 i + 15 of the group) meets, at step q, FDL stream row j = FILL - 1 - 16 i + q; its
// register window holds the 15 rows before that.  Chunks are issued in the order the consumers need them: the FILLC
// chunks of the fill, then filter chunk k / FDL chunk FILLC + k alternately.  Chunk c of a stream lives in ring slot
// c mod ring size, phase parity (c / ring size) & 1; "full" barriers count the producer's expect_tx + the copied bytes,
// "empty" barriers one arrival per consumer warp.
//
// The producer is not a warp of its own: a ninth warp would put five warps on one scheduler and cap the kernel at 96
// registers (the register file is per scheduler), which spills.  The last consumer warp issues the copies at the top
// of every group of 16 steps instead: what this group needs (blocking, normally long done) and up to LA chunks beyond
// (only while ring slots are free), so the copies run one to two groups ahead of the arithmetic, across segment
// boundaries.
template <int NTILE>
__global__ void __launch_bounds__(TbsCfg<NTILE>::THREADS, 2)
k_fdl_mac_tbs(const MacSeg* __restrict__ segs, const uint32_t* __restrict__ cta_seg_begin, uint32_t n_plan_ctas,
              uint32_t plan_per_cta, const float2* __restrict__ fdl, float2* __restrict__ ypart, uint32_t B, uint32_t R,
              uint32_t head0, uint32_t t0, uint32_t nt, uint32_t ncoltiles, uint32_t slot_stride, int* __restrict__ status) {
  using C = TbsCfg<NTILE>;
  constexpr int TT = C::TT, COLS = C::COLS, ROWB = C::ROWB, CH = C::CH, CHB = C::CHB, FILL = C::FILL, FILLC = C::FILLC;
  constexpr int XCH = C::XCH, HCH = C::HCH, NCONS = C::NCONS, GC = C::GC, LA = C::LA;
  extern __shared__ __align__(128) uint8_t tbs_smem[];
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(tbs_smem);
  const uint32_t xring = smem0, hring = smem0 + XCH * CHB;
  const uint32_t xfull = hring + HCH * CHB, xempty = xfull + 8 * XCH, hfull = xempty + 8 * XCH, hempty = hfull + 8 * HCH;

  const uint32_t coltile = blockIdx.x % ncoltiles, tgroup = blockIdx.x / ncoltiles;
  const uint32_t tbase0 = tgroup * (TT * NTILE);     // first block-step of this tile group, relative to t0
  const uint32_t s0 = (head0 + t0 + tbase0) % R;     // its FDL slot
  const uint32_t col0 = coltile * COLS;
  const uint32_t pc0 = blockIdx.y * plan_per_cta, pc1 = min(pc0 + plan_per_cta, n_plan_ctas);
  if (pc0 >= pc1) return;
  const uint32_t sb = cta_seg_begin[pc0], se = cta_seg_begin[pc1];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_prod = warp == NCONS - 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < XCH; s++) {
      ac::mbar_init(xfull + 8 * s, 1);
      ac::mbar_init(xempty + 8 * s, NCONS);
    }
    for (int s = 0; s < HCH; s++) {
      ac::mbar_init(hfull + 8 * s, 1);
      ac::mbar_init(hempty + 8 * s, NCONS);
    }
    ac::mbar_init_fence();
  }
  // a wait that times out (protocol bug) poisons the result but never hangs the device: the first time-out sets a word
  // in shared memory, after which every wait gives up after a handful of polls
  const uint32_t dead = smem0 + C::SMEM - 8;
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(tbs_smem + C::SMEM - 8) = 0u;
  __syncthreads();
  auto wait = [&](uint32_t bar, uint32_t parity) {
    if (!ac::mbar_try_wait(bar, parity)) ac::mbar_wait_or_die(bar, parity, dead);
  };

  // ---- producer cursor (meaningful in the producer warp only; uniform across its lanes) ----
  uint32_t p_si = sb;          // segment being issued
  uint32_t p_xk = 0, p_hk = 0;  // chunks of that segment issued so far
  uint32_t p_xn = 0, p_hn = 0;  // chunks issued since the kernel started = global index of the next chunk
  // issue the next chunk of the sequence; blocking = wait for its ring slot, else give up when the slot is still in use
  auto issue_next = [&](bool blocking) -> bool {
    const MacSeg sg = segs[p_si];
    const uint32_t np = sg.np, nx = FILL - 1 + np;
    const uint32_t nxc = (nx + CH - 1) / CH, nhc = (np + CH - 1) / CH;
    const bool want_x = p_xk < (uint32_t)FILLC || !(p_hk < nhc && p_hk + FILLC <= p_xk);
    if (want_x && p_xk < nxc) {
      const uint32_t slot = p_xn & (XCH - 1), par = ((p_xn / XCH) & 1u) ^ 1u;
      if (blocking) {
        wait(xempty + 8 * slot, par);
      } else if (!ac::mbar_try_wait(xempty + 8 * slot, par)) {
        return false;
      }
      const uint32_t rows = min((uint32_t)CH, nx - p_xk * CH);
      if (lane == 0) ac::mbar_expect_tx(xfull + 8 * slot, rows * ROWB);
      __syncwarp();
      if (lane < rows) {
        uint32_t base = s0 + R - (sg.p0 % R);
        if (base >= R) base -= R;
        int r = (int)(base + FILL - 1) - (int)(p_xk * CH + lane);
        r %= (int)R;
        if (r < 0) r += (int)R;
        const float2* src = fdl + ((uint64_t)sg.fdl_ch * R + (uint32_t)r) * B + col0;
        ac::bulk_g2s(xring + slot * CHB + lane * ROWB, src, ROWB, xfull + 8 * slot);
      }
      p_xk++;
      p_xn++;
    } else {
      const uint32_t slot = p_hn & (HCH - 1), par = ((p_hn / HCH) & 1u) ^ 1u;
      if (blocking) {
        wait(hempty + 8 * slot, par);
      } else if (!ac::mbar_try_wait(hempty + 8 * slot, par)) {
        return false;
      }
      const uint32_t rows = min((uint32_t)CH, np - p_hk * CH);
      if (lane == 0) ac::mbar_expect_tx(hfull + 8 * slot, rows * ROWB);
      __syncwarp();
      if (lane < rows) {
        const float2* src = reinterpret_cast<const float2*>(sg.H) + (uint64_t)(sg.p0 + p_hk * CH + lane) * B + col0;
        ac::bulk_g2s(hring + slot * CHB + lane * ROWB, src, ROWB, hfull + 8 * slot);
      }
      p_hk++;
      p_hn++;
    }
    if (p_xk >= nxc && p_hk >= nhc) {  // segment complete
      p_si++;
      p_xk = p_hk = 0;
    }
    return true;
  };
  // everything up to (need_x, need_h) chunks is issued when this returns; then up to LA chunks more, while slots are free
  auto produce = [&](uint32_t need_x, uint32_t need_h) {
    while (p_si < se && (p_xn < need_x || p_hn < need_h)) issue_next(true);
    while (p_si < se && (p_xn < need_x + LA || p_hn < need_h + LA))
      if (!issue_next(false)) break;
  };

  // ==================================== consumers: thread = (bin, tile of 16 block-steps) ==============================
  const uint32_t col = threadIdx.x % COLS, tile = threadIdx.x / COLS;
  const uint32_t xbase = xring + col * 8, hbase = hring + col * 8;
  // FDL chunk (relative to the segment's first) that holds this tile's row of step 0 -- its last row; tile 0 reads
  // tile * GC chunks ahead of that
  const uint32_t xfirst = (FILL - TT * tile) / CH - 1;
  float2 acc[TT], W[TT];
#pragma unroll
  for (int i = 0; i < TT; i++) acc[i] = make_float2(0.f, 0.f);
  uint32_t xc = 0, hc = 0;  // first chunk of the current segment's streams
  for (uint32_t si = sb; si < se; si++) {
    const MacSeg sg = segs[si];
    const uint32_t np = sg.np;
    if (sg.flags & 1u) {
#pragma unroll
      for (int i = 0; i < TT; i++) acc[i] = make_float2(0.f, 0.f);
    }
    // ---- the fill: FDL chunks 0 .. FILLC-1 of the segment hold every tile's register window and its row of step 0 ----
    if (is_prod) produce(xc + FILLC, hc);
#pragma unroll
    for (int k = 0; k < FILLC; k++) {
      const uint32_t c = xc + k;
      wait(xfull + 8 * (c & (XCH - 1)), (c / XCH) & 1u);
    }
    // W[e] = row base_i + e = stream row FILL - 1 - 16 i - e (e = 1 .. 15); W[0] is overwritten by step 0
    {
      const uint32_t j0 = FILL - TT * (tile + 1);  // stream row of e = 15
#pragma unroll
      for (int e = 1; e < TT; e++) {
        const uint32_t j = j0 + (TT - 1 - e);
        const uint32_t c = xc + j / CH;
        W[e] = ac::lds2(xbase + (c & (XCH - 1)) * CHB + (j % CH) * ROWB);
      }
      W[0] = make_float2(0.f, 0.f);
    }
    // the chunks before the one that holds my row of step 0 are done (never read by this tile, or read by the fill)
    __syncwarp();
    if (lane == 0)
      for (uint32_t c = xc; c < xc + xfirst; c++) ac::mbar_arrive(xempty + 8 * (c & (XCH - 1)));
    uint32_t xcur = xbase + ((xc + xfirst) & (XCH - 1)) * CHB;
    uint32_t hcur = hbase;

    // q + u = step of the segment; u = (q + u) mod 16 because groups start at multiples of 16, so chunk boundaries fall
    // on fixed u.  Filter chunk of the step: hc + (q + u) / CH.  My FDL chunk: xc + xfirst + (q + u + CH - 1) / CH.
    // A chunk is handed back when the warp moves on to the next one (its last row went through the FMAs a step earlier).
    auto step = [&](const int u, const uint32_t q) {
      if (u % CH == 0) {
        const uint32_t hk = hc + q / CH + u / CH;
        if (q + u != 0) {
          __syncwarp();
          if (lane == 0) ac::mbar_arrive(hempty + 8 * ((hk - 1) & (HCH - 1)));
        }
        hcur = hbase + (hk & (HCH - 1)) * CHB;
        wait(hfull + 8 * (hk & (HCH - 1)), (hk / HCH) & 1u);
      }
      if (u % CH == 1) {
        const uint32_t xk = xc + xfirst + q / CH + u / CH + 1;
        __syncwarp();
        if (lane == 0) ac::mbar_arrive(xempty + 8 * ((xk - 1) & (XCH - 1)));
        xcur = xbase + (xk & (XCH - 1)) * CHB;
        const uint32_t w = xk + tile * GC;  // tile 0's chunk: the newest one any tile touches
        wait(xfull + 8 * (w & (XCH - 1)), (w / XCH) & 1u);
      }
      const float2 h = ac::lds2(hcur + (u % CH) * ROWB);
      W[(TT - u) % TT] = ac::lds2(xcur + ((u + CH - 1) % CH) * ROWB);
#pragma unroll
      for (int i = 0; i < TT; i++) cmac_x2(acc[i], h, W[(i - u + TT) % TT]);
    };

    uint32_t q = 0;
    for (; q + TT <= np; q += TT) {
      // chunks the group q .. q + 15 reads: filter chunks up to q / CH + GC, FDL chunks up to FILLC + q / CH + GC
      if (is_prod) {
        const uint32_t nxc = (FILL - 1 + np + CH - 1) / CH, nhc = (np + CH - 1) / CH;
        produce(xc + min(nxc, FILLC + q / CH + GC), hc + min(nhc, q / CH + GC));
      }
#pragma unroll
      for (int u = 0; u < TT; u++) step(u, q);
    }
    const uint32_t nxc = (FILL - 1 + np + CH - 1) / CH, nhc = (np + CH - 1) / CH;
    if (q < np) {
      if (is_prod) produce(xc + nxc, hc + nhc);
#pragma unroll
      for (int u = 0; u < TT; u++)
        if (q + u < np) step(u, q);
    }
    // chunks of this segment the warp has not handed back yet: the ones it was still reading, short last chunks, rows
    // only the other tiles read
    __syncwarp();
    if (lane == 0) {
      for (uint32_t c = xc + xfirst + (np - 1 + CH - 1) / CH; c < xc + nxc; c++) ac::mbar_arrive(xempty + 8 * (c & (XCH - 1)));
      for (uint32_t c = hc + (np - 1) / CH; c < hc + nhc; c++) ac::mbar_arrive(hempty + 8 * (c & (HCH - 1)));
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
  if (lane == 0 && status && *reinterpret_cast<volatile uint32_t*>(tbs_smem + C::SMEM - 8)) *status = 2;
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

template <int NTILE>
static cudaError_t launch_tbs_t(const MacTbsArgs& a, cudaStream_t st) {
  using C = TbsCfg<NTILE>;
  static uint32_t attr_set = 0;  // per device: function attributes belong to the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set & (1u << (dev & 31)))) {
    cudaError_t e = cudaFuncSetAttribute(k_fdl_mac_tbs<NTILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set |= 1u << (dev & 31);
  }
  const uint32_t ncol = a.B / C::COLS, ngroups = ceil_div(a.nt, (uint32_t)(C::TT * NTILE));
  // persistent: about two CTAs per SM in total, each walking consecutive row ranges of the plan
  uint32_t gy = std::max(1u, (2u * kNumSMs) / (ncol * ngroups));
  gy = std::min(gy, a.n_plan_ctas);
  const uint32_t per = ceil_div(a.n_plan_ctas, gy);
  gy = ceil_div(a.n_plan_ctas, per);
  k_fdl_mac_tbs<NTILE><<<dim3(ncol * ngroups, gy), C::THREADS, C::SMEM, st>>>(a.segs, a.cta_seg_begin, a.n_plan_ctas, per, a.fdl, a.ypart,
                                                                             a.B, a.R, a.head, a.t0, a.nt, ncol, a.slot_stride, a.status);
  return cudaGetLastError();
}

cudaError_t launch_nyq_mac2(const MacTbsArgs& a, cudaStream_t st) {
  k_nyq_mac2<<<dim3(a.n_plan_ctas, ceil_div(a.nt, 64u)), 64, 0, st>>>(a.segs, a.cta_seg_begin, a.fdl, a.nyq_part, a.B, a.R, a.head, a.t0,
                                                                      a.nt, a.slot_stride);
  return cudaGetLastError();
}

cudaError_t launch_mac_tbs(const MacTbsArgs& a, cudaStream_t st, const char** kernel_name) {
  // tiles per CTA: four (64 block-steps per CTA) unless the call is short; the column tile must fit the block size
  const bool four = a.nt > 32 || a.B < 128;
  if (four) {
    *kernel_name = "k_fdl_mac_tbs<4>";
    return launch_tbs_t<4>(a, st);
  }
  *kernel_name = "k_fdl_mac_tbs<2>";
  return launch_tbs_t<2>(a, st);
}

}  // namespace bbx
