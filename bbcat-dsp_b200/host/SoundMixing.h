// MixSamples with the reference's signatures (src/SoundMixing.h:55-106), executed by libbbx.
#pragma once

#include "Interpolator.h"
#include "SoundFormatConversions.h"

namespace bbcat {

inline void MixSamples(const float* src, uint_t src_channel, uint_t src_channels, float* dst, uint_t dst_channel,
                       uint_t dst_channels, uint_t nchannels, uint_t nframes, float mul = 1.0f) {
  (void)bbx_mix_samples_f32(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul);
}

inline void MixSamples(const double* src, uint_t src_channel, uint_t src_channels, double* dst, uint_t dst_channel,
                       uint_t dst_channels, uint_t nchannels, uint_t nframes, double mul = 1.0) {
  (void)bbx_mix_samples_f64(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul);
}

// level ramps with the caller's Interpolator, which is advanced nframes steps
inline void MixSamples(const Sample_t* src, uint_t src_channel, uint_t src_channels, Sample_t* dst, uint_t dst_channel,
                       uint_t dst_channels, uint_t nchannels, uint_t nframes, Interpolator& interp, Sample_t inc) {
  (void)bbx_mix_samples_interp(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes,
                               interp.State(), inc);
}

}  // namespace bbcat
