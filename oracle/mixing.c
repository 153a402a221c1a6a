/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of
 *   MixSamples<T>                              SoundMixing.h:55-81
 *   MixSamples(..., Interpolator&, inc)        SoundMixing.cpp:23-52
 *   Interpolator::operator+= / NonZero         Interpolator.h:25,55
 * Arithmetic: one rounding of the product, one of the sum (the reference is built
 * with -msse3, no FMA; this file is built with -ffp-contract=off).
 */
#include "oracle.h"

void orc_mix_samples_f32(const float* src, unsigned src_channel, unsigned src_channels, float* dst,
                         unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes, float mul) {
  unsigned i, j;
  if (!orc_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 1))
    return;
  if (!(mul != 0.0f)) return; /* (mul != T()) : a zero gain is a no-op, a NaN gain is not */
  src += src_channel;
  dst += dst_channel;
  for (i = 0; i < nframes; i++, src += src_channels, dst += dst_channels)
    for (j = 0; j < nchannels; j++) {
      float prod = mul * src[j];
      dst[j] = dst[j] + prod;
    }
}

void orc_mix_samples_f64(const double* src, unsigned src_channel, unsigned src_channels, double* dst,
                         unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes,
                         double mul) {
  unsigned i, j;
  if (!orc_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 1))
    return;
  if (!(mul != 0.0)) return;
  src += src_channel;
  dst += dst_channel;
  for (i = 0; i < nframes; i++, src += src_channels, dst += dst_channels)
    for (j = 0; j < nchannels; j++) {
      double prod = mul * src[j];
      dst[j] = dst[j] + prod;
    }
}

/* Interpolator.h:55 : step current towards target by inc, never past it */
static float interp_step(float target, float current, float inc) {
  if (target >= current) {
    float v = current + inc;
    return (target < v) ? target : v; /* std::min(current + inc, target) */
  } else {
    float v = current - inc;
    return (v < target) ? target : v; /* std::max(current - inc, target) */
  }
}

void orc_interpolator_step(float* st, float inc, unsigned nsteps) {
  unsigned i;
  for (i = 0; i < nsteps; i++) st[1] = interp_step(st[0], st[1], inc);
}

void orc_mix_samples_interp(const float* src, unsigned src_channel, unsigned src_channels, float* dst,
                            unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes,
                            float* st, float inc) {
  unsigned i, j;
  /* no single-frame collapse here: the gain changes per frame (SoundMixing.cpp:32-36) */
  if (!orc_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 0))
    return;
  if (!((st[1] != 0.0f) || (st[0] != 0.0f))) return; /* interp.NonZero() */
  src += src_channel;
  dst += dst_channel;
  float mul = st[1];
  for (i = 0; i < nframes; i++, src += src_channels, dst += dst_channels) {
    for (j = 0; j < nchannels; j++) {
      float prod = mul * src[j];
      dst[j] = dst[j] + prod;
    }
    st[1] = interp_step(st[0], st[1], inc);
    mul = st[1];
  }
}
