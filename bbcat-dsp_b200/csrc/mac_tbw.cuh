// mac_tbw.cuh -- k_fdl_mac_tbw: the time-batched FDL MAC with WARP-PRIVATE operand streams.
//
// Same sums, plan and FMA order as the other MAC kernels (bit-identical).  Thread = (bin, tile of 16 block-steps) as in
// k_fdl_mac_tbs, but the unit that shares filter and FDL rows is one warp: 8 bins x 4 time tiles.  Every warp streams its
// own 64-byte row pieces into its own shared-memory rings with 16-byte cp.async copies -- a lane takes one piece of two
// rows 8 apart per group of 16 rows and stream, so that its cursors (global pointer, ring row, shared-memory target) move
// once per group -- and tracks them with cp.async groups (commit / wait_group) and __syncwarp only: no block barrier, no
// mbarrier.  The shared-memory addresses of the operands a thread reads advance by one group per iteration as well.  The warps
// of a CTA walk different row ranges of the plan, so their segment boundaries (where a fill has to be waited for) fall at
// different times and the other warps of the scheduler keep the FMA pipe busy meanwhile.
//   * k_fdl_mac_tbs (block-shared rings) needs a block barrier per group of 16 steps; it ends up at the speed of round
//     1's kernel (0.217 ms against 0.213): the barriers put all warps of a CTA in phase, so their non-FMA stretches --
//     barrier skew, copy issue, waiting for a fill -- coincide.
//   * the four tiles of a warp read FDL rows 16 apart in the same instruction (32 lanes x 8 bytes = 2 wavefronts at
//     best).  Rows are 64 bytes = 16 banks wide, so rows a multiple of 16 apart would all sit in the same half of the
//     banks (4 wavefronts); one 64-byte pad per 16 rows makes the halves alternate.
#pragma once

#include "async_copy.cuh"
#include "mac_common.cuh"

namespace bbx {

struct TbwCfg {
  static constexpr int TT = 16, NTILE = 4, COLS = 8, ROWB = 64, CH = 8;
  static constexpr int FILL = TT * NTILE;          // 64 FDL rows before a segment's first step
  static constexpr int XR = 128, HR = 64;          // ring rows (powers of two): live window 64 + two groups ahead
  static constexpr int XGROUP_BYTES = 17 * ROWB;   // 16 rows + one pad row
  static constexpr int XBYTES = (XR / 16) * XGROUP_BYTES, HBYTES = HR * ROWB;
  static constexpr int WARP_SMEM = XBYTES + HBYTES;  // 12800 bytes
  static constexpr int WARPS = 8;
  static constexpr int SMEM = WARPS * WARP_SMEM;     // 102400 bytes per CTA, two CTAs per SM
  static constexpr int LOOK = 2;                     // rounds (groups of 16 steps) copied ahead of the one being computed
};

// Register budget: two CTAs of 256 threads per SM.  The register file is allocated in units of 8 per thread, and with 121 - 128
// registers two CTAs would need all 64 K of it: two unrelated builds in that range both took 0.2068 ms per C3 launch, which
// is what one CTA per SM in two waves would give.  Below that, fewer is better down to 112 (0.1833 ms at 116, 0.1806 at 112 for the same source; 104 and 96
// spill: 0.187 / 0.195 ms), so the budget is stated instead of left to __launch_bounds__.
#ifndef BBX_TBW_MAXNREG
#define BBX_TBW_MAXNREG 112
#endif
__global__ void __maxnreg__(BBX_TBW_MAXNREG)
k_fdl_mac_tbw(const MacSeg* __restrict__ segs, const uint32_t* __restrict__ cta_seg_begin, uint32_t n_plan_ctas,
              uint32_t plan_per_warp, uint32_t n_ranges, const float2* __restrict__ fdl, float2* __restrict__ ypart, uint32_t B,
              uint32_t R, uint32_t head0, uint32_t t0, uint32_t nt, uint32_t ncoltiles, uint32_t ngroups, uint32_t slot_stride) {
  using C = TbwCfg;
  constexpr int TT = C::TT, CH = C::CH, FILL = C::FILL, ROWB = C::ROWB, LOOK = C::LOOK;
  extern __shared__ __align__(128) uint8_t tbw_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warp -> (row range, column tile, tile group): consecutive warps take different row ranges
  const uint32_t wid = blockIdx.x * C::WARPS + warp;
  const uint32_t range = wid % n_ranges, rest = wid / n_ranges;
  const uint32_t coltile = rest % ncoltiles, tgroup = rest / ncoltiles;
  if (tgroup >= ngroups) return;
  const uint32_t pc0 = range * plan_per_warp, pc1 = min(pc0 + plan_per_warp, n_plan_ctas);
  if (pc0 >= pc1) return;
  const uint32_t sb = cta_seg_begin[pc0], se = cta_seg_begin[pc1];
  const uint32_t xsm = (uint32_t)__cvta_generic_to_shared(tbw_smem) + warp * C::WARP_SMEM, hsm = xsm + C::XBYTES;
  const uint32_t tbase0 = tgroup * (TT * C::NTILE);
  const uint32_t s0 = (head0 + t0 + tbase0) % R;
  const uint32_t col0 = coltile * C::COLS;
  const uint32_t tile = lane >> 3, col = lane & 7;
  // copies: lane = (row of the chunk, 16-byte piece of the row)
  const uint32_t crow = lane >> 2, cpiece = (lane & 3) * 16;
  const uint32_t chunk_bytes_g = CH * B * (uint32_t)sizeof(float2);
  const uint64_t ring_bytes_g = (uint64_t)R * B * sizeof(float2);
  const uint64_t ystride = (uint64_t)slot_stride * B;  // float2 elements between the partial sums of consecutive block-steps

  float2 acc[TT], W[TT];
#pragma unroll
  for (int i = 0; i < TT; i++) acc[i] = make_float2(0.f, 0.f);
  uint32_t xrow0 = 0, hrow0 = 0;  // ring rows (multiples of 16) where the current segment's streams start
  for (uint32_t si = sb; si < se; si++) {
    const MacSeg sg = segs[si];
    const uint32_t np = sg.np, nx = FILL - 1 + np;
    if (sg.flags & 1u) {
#pragma unroll
      for (int i = 0; i < TT; i++) acc[i] = make_float2(0.f, 0.f);
    }
    // ---- this lane's copy cursors for the segment: two rows 8 apart per copy step (one cursor update per 16 rows) ----
    uint32_t base = s0 + R - sg.p0;  // p0 < P <= R (R = P + T_max - 1): no reduction of p0 needed
    if (base >= R) base -= R;
    int xr = (int)(base + FILL - 1) - (int)crow;  // FDL ring row of my first row of the next pair
    while (xr >= (int)R) xr -= (int)R;
    const char* xptr = reinterpret_cast<const char*>(fdl + ((uint64_t)sg.fdl_ch * R + (uint32_t)xr) * B + col0) + cpiece;
    const char* hptr = reinterpret_cast<const char*>(reinterpret_cast<const float2*>(sg.H) + (uint64_t)(sg.p0 + crow) * B + col0) + cpiece;
    uint32_t xj = crow, hq = crow;  // stream row of my first row of the next pair
    // shared-memory targets of my first row: the X ring advances by one padded group of 16 rows per pair, the H ring by 16 rows
    uint32_t xdst, hoff;
    {
      const uint32_t rho = (xrow0 + crow) & (C::XR - 1);
      xdst = xsm + (rho + (rho >> 4)) * ROWB + cpiece;
      hoff = ((hrow0 + crow) & (C::HR - 1)) * ROWB;
    }
    auto copy_x_pair = [&]() {
      if (xj < nx) ac::cp_async16(xdst, xptr);
      if (xj + CH < nx) {
        const char* x2 = xptr - chunk_bytes_g;  // eight rows further down the stream = eight ring rows lower
        if (xr < CH) x2 += ring_bytes_g;
        ac::cp_async16(xdst + CH * ROWB, x2);
      }
      xj += 2 * CH;
      xdst += C::XGROUP_BYTES;
      if (xdst >= xsm + C::XBYTES) xdst -= C::XBYTES;
      xr -= 2 * CH;
      xptr -= 2 * (uint64_t)chunk_bytes_g;
      if (xr < 0) {
        xr += (int)R;
        xptr += ring_bytes_g;
      }
    };
    auto copy_h_pair = [&]() {
      const uint32_t dst = hsm + hoff + cpiece;
      if (hq < np) ac::cp_async16(dst, hptr);
      if (hq + CH < np) ac::cp_async16(dst + CH * ROWB, hptr + chunk_bytes_g);
      hq += 2 * CH;
      hoff = (hoff + 2 * CH * ROWB) & (C::HBYTES - 1);
      hptr += 2 * (uint64_t)chunk_bytes_g;
    };
    // round 0 = the fill + the first group's rows; round g = group g's rows (16 of each stream)
    auto copy_round = [&](bool first) {
      if (first) {
#pragma unroll
        for (int k = 0; k < FILL / (2 * CH); k++) copy_x_pair();
      }
      copy_x_pair();
      copy_h_pair();
      ac::cp_async_commit();
    };
    __syncwarp();  // every lane is through with the previous segment's rows before they are overwritten
    copy_round(true);
#pragma unroll
    for (int r = 1; r < LOOK; r++) copy_round(false);

    // ---- compute: q + u = step of the segment, groups of 16 start at multiples of 16 (and so do the ring rows) ----
    // my tile's FDL row of step q + u sits at ring row (xrow0 + q + 63 - 16 tile + u): low four bits 15 (u = 0) or u - 1;
    // the two groups and the filter rows advance by one group per iteration
    uint32_t xb0, xb1, hb, hboff = (hrow0 & (C::HR - 1)) * ROWB;
    {
      const uint32_t g0 = (xrow0 / 16 + 3 - tile) & (C::XR / 16 - 1), g1 = (g0 + 1) & (C::XR / 16 - 1);
      xb0 = xsm + g0 * C::XGROUP_BYTES + col * 8;
      xb1 = xsm + g1 * C::XGROUP_BYTES + col * 8;
      hb = hsm + hboff + col * 8;
    }
    float2 hn = make_float2(0.f, 0.f), xn = hn;
    auto fetch = [&](const int u) {
      hn = ac::lds2(hb + u * ROWB);
      xn = ac::lds2((u == 0 ? xb0 + 15 * ROWB : xb1 + (u - 1) * ROWB));
    };
    auto step = [&](const int u, const bool more) {
      const float2 h = hn;
      W[(TT - u) % TT] = xn;
      if (more && u + 1 < TT) fetch(u + 1);
#pragma unroll
      for (int i = 0; i < TT; i++) cmac_x2(acc[i], h, W[(i - u + TT) % TT]);
    };
    for (uint32_t q = 0; q < np; q += TT) {
      copy_round(false);           // round q / 16 + LOOK
      ac::cp_async_wait_group<LOOK>();  // round q / 16 has landed (my copies)
      __syncwarp();                // ... and everybody else's
      fetch(0);
      if (q == 0) {
        // W[e] = FDL stream row 63 - 16 tile - e (e = 1 .. 15): ring rows of group g0, low bits 15 - e
#pragma unroll
        for (int e = 1; e < TT; e++) W[e] = ac::lds2(xb0 + (15 - e) * ROWB);
        W[0] = make_float2(0.f, 0.f);
      }
      if (q + TT <= np) {
#pragma unroll
        for (int u = 0; u < TT; u++) step(u, true);
      } else {
#pragma unroll
        for (int u = 0; u < TT; u++)
          if (q + u < np) step(u, q + u + 1 < np);
      }
      xb0 = xb1;
      xb1 += C::XGROUP_BYTES;
      if (xb1 >= xsm + C::XBYTES + col * 8) xb1 -= C::XBYTES;
      hboff = (hboff + TT * ROWB) & (C::HBYTES - 1);
      hb = hsm + hboff + col * 8;
    }
    xrow0 = (xrow0 + ((nx + 15) & ~15u)) & (C::XR - 1);
    hrow0 = (hrow0 + ((np + 15) & ~15u)) & (C::HR - 1);
    if (sg.flags & 2u) {
      // one address, then a constant stride per block-step; full tiles (every tile of a 64-block call) skip the bound checks
      const uint32_t tb = tbase0 + tile * TT;
      float2* yp = ypart + ((uint64_t)tb * slot_stride + sg.slot) * B + col0 + col;
      if (tb + TT <= nt) {
#pragma unroll
        for (int i = 0; i < TT; i++, yp += ystride) *yp = acc[i];
      } else {
#pragma unroll
        for (int i = 0; i < TT; i++, yp += ystride)
          if (tb + i < nt) *yp = acc[i];
      }
    }
  }
  ac::cp_async_wait_all();
}

}  // namespace bbx
