/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of FractionalSample (float and double overloads),
 * FractionalSample.cpp:249-341: 14-tap x 128-phase polyphase read from a circular,
 * possibly interleaved buffer, accumulated in double, taps in ascending order.
 * Coefficients: fracsample_table.inc (data re-encoded from FractionalSample.cpp:17-243).
 */
#include "oracle.h"

enum { OVERSAMPLE = 128, TAPS = 14 };

static const double k_filter[OVERSAMPLE * TAPS] = {
#include "fracsample_table.inc"
};

unsigned orc_fractional_sample_additional_delay_required(void) { return TAPS; }

#define FRAC_BODY(T)                                                                              \
  unsigned fpos = OVERSAMPLE - 1 - ((unsigned)((double)OVERSAMPLE * pos) % OVERSAMPLE);           \
  unsigned bpos = (unsigned)pos + length - TAPS;                                                  \
  double res = 0.0;                                                                               \
  int t;                                                                                          \
  buffer += channel;                                                                              \
  bpos *= channels;                                                                               \
  length *= channels;                                                                             \
  bpos %= length;                                                                                 \
  for (t = 0; t < TAPS; t++) {                                                                    \
    double prod = k_filter[fpos] * (double)buffer[bpos];                                          \
    res = res + prod;                                                                             \
    fpos += OVERSAMPLE;                                                                           \
    bpos += channels;                                                                             \
    if (bpos >= length) bpos -= length;                                                           \
  }                                                                                               \
  return res;

double orc_fractional_sample_f32(const float* buffer, unsigned channel, unsigned channels, unsigned length, double pos) {
  FRAC_BODY(float)
}

double orc_fractional_sample_f64(const double* buffer, unsigned channel, unsigned channels, unsigned length, double pos) {
  FRAC_BODY(double)
}

void orc_fractional_samples_f32(const float* buffer, unsigned channel, unsigned channels, unsigned length,
                                const double* pos, unsigned n, double* out) {
  unsigned i;
  for (i = 0; i < n; i++) out[i] = orc_fractional_sample_f32(buffer, channel, channels, length, pos[i]);
}

void orc_fractional_samples_f64(const double* buffer, unsigned channel, unsigned channels, unsigned length,
                                const double* pos, unsigned n, double* out) {
  unsigned i;
  for (i = 0; i < n; i++) out[i] = orc_fractional_sample_f64(buffer, channel, channels, length, pos[i]);
}
