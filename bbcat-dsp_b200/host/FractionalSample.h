// FractionalSample with the reference's signatures (src/FractionalSample.h:11-38).  The scalar form is
// kept for drop-in use; callers that read whole blocks should use FractionalSamples() (one launch).
#pragma once

#include "SoundFormatConversions.h"

namespace bbcat {

inline uint_t FractionalSampleAdditionalDelayRequired() { return bbx_fractional_sample_additional_delay_required(); }

inline void FractionalSamples(const float* buffer, uint_t channel, uint_t channels, uint_t length, const double* pos, uint_t n,
                              double* out) {
  (void)bbx_fractional_samples_f32(buffer, channel, channels, length, pos, n, out);
}
inline void FractionalSamples(const double* buffer, uint_t channel, uint_t channels, uint_t length, const double* pos, uint_t n,
                              double* out) {
  (void)bbx_fractional_samples_f64(buffer, channel, channels, length, pos, n, out);
}

inline double FractionalSample(const float* buffer, uint_t channel, uint_t channels, uint_t length, double pos) {
  double res = 0.0;
  FractionalSamples(buffer, channel, channels, length, &pos, 1, &res);
  return res;
}
inline double FractionalSample(const double* buffer, uint_t channel, uint_t channels, uint_t length, double pos) {
  double res = 0.0;
  FractionalSamples(buffer, channel, channels, length, &pos, 1, &res);
  return res;
}

}  // namespace bbcat
