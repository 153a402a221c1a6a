// fracsample.cuh -- FractionalSample on the device (bbcat-dsp src/FractionalSample.cpp:249-341).
//
// 14-tap x 128-phase polyphase read from a circular, possibly interleaved buffer.  The phase is
// floor-quantised to 1/128 sample, the filter starts 14 frames back from `pos`, taps run in
// ascending order and accumulate in double with the product and the sum rounded separately, so the
// result equals the reference's x86 (no-FMA) build bit for bit.
// The coefficient table is data re-encoded from FractionalSample.cpp:17-243 (see the .inc header).
#pragma once

#include <stdint.h>

namespace bbx {

static __device__ const double g_frac_filter[128 * 14] = {
#include "fracsample_table.inc"
};

template <typename T>
__device__ __forceinline__ double fractional_sample_dev(const T* __restrict__ buffer, uint32_t channel, uint32_t channels,
                                                        uint32_t length, double pos) {
  uint32_t fpos = 127u - ((uint32_t)(128.0 * pos) % 128u);  // .cpp:314
  uint32_t bpos = (uint32_t)pos + length - 14u;             // .cpp:315
  double res = 0.0;
  buffer += channel;
  bpos *= channels;
  length *= channels;
  bpos %= length;
#pragma unroll
  for (int t = 0; t < 14; t++) {
    res = __dadd_rn(res, __dmul_rn(g_frac_filter[fpos], (double)buffer[bpos]));
    fpos += 128u;
    bpos += channels;
    if (bpos >= length) bpos -= length;
  }
  return res;
}

}  // namespace bbx
