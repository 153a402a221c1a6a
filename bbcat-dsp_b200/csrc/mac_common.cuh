// mac_common.cuh -- the MAC plan segment and the one complex multiply-accumulate every MAC kernel of the engine uses.
#pragma once

#include "kernels_common.cuh"

namespace bbx {

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

}  // namespace bbx
