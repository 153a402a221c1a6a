/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * biquad.c -- C restatement of BiQuadCoeffs / BiQuad ("next" row):
 *   coefficient design      src/BiQuad.cpp:181-352  (CalcCoeffs; filter types src/BiQuad.h:31-42)
 *   explicit coefficients   src/BiQuad.cpp:75-103   (SetCoeffs, interpolation time in samples)
 *   coefficient ramp        src/BiQuad.cpp:379-395  (Interpolate: mul -= dec, current = target - mul * diff)
 *   the filter              src/BiQuad.h:200-206    (direct form II transposed, double state, float in/out)
 *   multi-channel process   src/BiQuad.cpp:473-497  (frame-major, one ramp step per frame)
 * Pinned against the reference's own BiQuad.cpp compiled into oracle/_ref (tests/test_biquad.py) and against
 * tests/golden/biquad.npz.  Compiled with -ffp-contract=off: every product and sum is rounded separately,
 * like the reference's SSE2 build.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

struct orc_biquad {
  unsigned nch;
  double cur[5], tgt[5], dif[5]; /* num0 num1 num2 den1 den2 */
  double mul, dec;
  double* w; /* [nch][2] */
};

/* target coefficients of a filter description, normalised by a0 (src/BiQuad.cpp:181-330) */
static void design(int type, double freq, double fs, double gain, double bandwidth, double* t) {
  const double A = pow(10.0, gain / 40.0);
  const double omega = 2.0 * M_PI * freq / fs;
  const double sn = sin(omega), cs = cos(omega);
  const double alpha = sn * sinh(M_LN2 / 2.0 * bandwidth * omega / sn);
  const double beta = sqrt(A + A);
  double b0 = 1.0, b1 = 0.0, b2 = 0.0, a0 = 1.0, a1 = 0.0, a2 = 0.0; /* FLAT and unknown types */
  switch (type) {
    case 1: /* LPF6 */
      b0 = sn; b1 = 0; b2 = 0; a0 = 1 + sn; a1 = -1; a2 = 0;
      break;
    case 3: /* LPF12 */
      b0 = sn * sn; b1 = 0; b2 = 0; a0 = (1 + sn) * (1 + sn); a1 = -2 * (1 + sn); a2 = 1;
      break;
    case 2: /* HPF6 */
      b0 = 1; b1 = -1; b2 = 0; a0 = 1; a1 = -(1 - sn); a2 = 0;
      break;
    case 4: /* HPF12 */
      b0 = 1; b1 = -2; b2 = 1; a0 = 1; a1 = -2 * (1 - sn); a2 = (1 - sn) * (1 - sn);
      break;
    case 5: /* BPF */
      b0 = alpha; b1 = 0; b2 = -alpha; a0 = 1 + alpha; a1 = -2 * cs; a2 = 1 - alpha;
      break;
    case 6: /* NOTCH */
      b0 = 1; b1 = -2 * cs; b2 = 1; a0 = 1 + alpha; a1 = -2 * cs; a2 = 1 - alpha;
      break;
    case 7: /* PEQ */
      b0 = 1 + (alpha * A); b1 = -2 * cs; b2 = 1 - (alpha * A);
      a0 = 1 + (alpha / A); a1 = -2 * cs; a2 = 1 - (alpha / A);
      break;
    case 8: /* LSH */
      b0 = A * ((A + 1) - (A - 1) * cs + beta * sn);
      b1 = 2 * A * ((A - 1) - (A + 1) * cs);
      b2 = A * ((A + 1) - (A - 1) * cs - beta * sn);
      a0 = (A + 1) + (A - 1) * cs + beta * sn;
      a1 = -2 * ((A - 1) + (A + 1) * cs);
      a2 = (A + 1) + (A - 1) * cs - beta * sn;
      break;
    case 9: /* HSH */
      b0 = A * ((A + 1) + (A - 1) * cs + beta * sn);
      b1 = -2 * A * ((A - 1) + (A + 1) * cs);
      b2 = A * ((A + 1) + (A - 1) * cs - beta * sn);
      a0 = (A + 1) - (A - 1) * cs + beta * sn;
      a1 = 2 * ((A - 1) - (A + 1) * cs);
      a2 = (A + 1) - (A - 1) * cs - beta * sn;
      break;
    default:
      break;
  }
  const double normalise = 1.0 / a0;
  t[0] = b0 * normalise;
  t[1] = b1 * normalise;
  t[2] = b2 * normalise;
  t[3] = a1 * normalise;
  t[4] = a2 * normalise;
}

/* new targets are in b->tgt: differences, then either start a ramp of `steps` samples or jump */
static void retarget(orc_biquad* b, double steps) {
  for (int i = 0; i < 5; i++) b->dif[i] = b->tgt[i] - b->cur[i];
  if (steps > 0.0) {
    b->mul = 1.0;
    b->dec = 1.0 / steps;
  } else {
    b->mul = b->dec = 0.0;
    memcpy(b->cur, b->tgt, sizeof(b->cur));
  }
}

void orc_biquad_calc_coeffs(int type, double freq, double fs, double gain, double bandwidth, double* out5) {
  design(type, freq, fs, gain, bandwidth, out5);
}

orc_biquad* orc_biquad_create(unsigned nch) {
  orc_biquad* b = (orc_biquad*)calloc(1, sizeof(*b));
  b->nch = nch;
  b->cur[0] = b->tgt[0] = 1.0; /* BiQuadCoeffs(): num0 = 1, the rest 0; mul 0, dec 1 (src/BiQuad.cpp:11-24) */
  b->dec = 1.0;
  b->w = (double*)calloc(2 * (size_t)(nch ? nch : 1), sizeof(double));
  return b;
}

void orc_biquad_destroy(orc_biquad* b) {
  if (!b) return;
  free(b->w);
  free(b);
}

void orc_biquad_set_coeffs(orc_biquad* b, const double* c5, double interp_samples) {
  memcpy(b->tgt, c5, sizeof(b->tgt));
  retarget(b, interp_samples);
}

void orc_biquad_calc(orc_biquad* b, int type, double freq, double fs, double gain, double bandwidth, double interp_time) {
  design(type, freq, fs, gain, bandwidth, b->tgt);
  retarget(b, interp_time > 0.0 ? interp_time * fs : 0.0);
}

void orc_biquad_process(orc_biquad* b, const float* src, float* dst, unsigned nchannels, unsigned nsrc, unsigned ndst,
                        unsigned nframes) {
  if (nchannels > nsrc) nchannels = nsrc;
  if (nchannels > ndst) nchannels = ndst;
  if (nchannels > b->nch) nchannels = b->nch;
  for (unsigned i = 0; i < nframes; i++, src += nsrc, dst += ndst) {
    for (unsigned j = 0; j < nchannels; j++) {
      double* w = b->w + 2 * (size_t)j;
      const float x = src[j];
      const float y = (float)(x * b->cur[0] + w[0]);
      w[0] = x * b->cur[1] - y * b->cur[3] + w[1];
      w[1] = x * b->cur[2] - y * b->cur[4];
      dst[j] = y;
    }
    if (b->mul > 0.0) { /* one ramp step per frame (src/BiQuad.cpp:379-395 with count = 1) */
      b->mul -= b->dec * 1.0;
      b->mul = b->mul > 0.0 ? b->mul : 0.0;
      for (int k = 0; k < 5; k++) b->cur[k] = b->tgt[k] - b->mul * b->dif[k];
    }
  }
}

void orc_biquad_get_state(const orc_biquad* b, double* w, double* cur5, double* mul_dec) {
  memcpy(w, b->w, 2 * (size_t)b->nch * sizeof(double));
  memcpy(cur5, b->cur, sizeof(b->cur));
  mul_dec[0] = b->mul;
  mul_dec[1] = b->dec;
}

void orc_biquad_reset(orc_biquad* b) { memset(b->w, 0, 2 * (size_t)b->nch * sizeof(double)); }

/* ---- BiQuadFilterBank (src/BiQuad.h:247-353, src/BiQuad.cpp:498-662) ----
 * nfilters coefficient objects, each with one filter state per channel.  Process is the reference's loop as written
 * (src/BiQuad.cpp:639-662): filter 0 from src to dst, every later filter in place on dst, each a full pass of
 * BiQuad::Process with that filter's own ramp.  SetFilters / SetChannels (src/BiQuad.cpp:528-600) drop or append at the end;
 * surviving (filter, channel) pairs keep their state, new ones start flat and silent. */
struct orc_fbank {
  unsigned nch, nfilters;
  orc_biquad** f;
};

orc_fbank* orc_fbank_create(unsigned nch, unsigned nfilters) {
  orc_fbank* b = (orc_fbank*)calloc(1, sizeof(*b));
  b->nch = nch;
  orc_fbank_set_filters(b, nfilters);
  return b;
}

void orc_fbank_destroy(orc_fbank* b) {
  if (!b) return;
  orc_fbank_set_filters(b, 0);
  free(b->f);
  free(b);
}

void orc_fbank_set_filters(orc_fbank* b, unsigned n) {
  for (unsigned i = n; i < b->nfilters; i++) orc_biquad_destroy(b->f[i]);
  b->f = (orc_biquad**)realloc(b->f, sizeof(*b->f) * (n ? n : 1));
  for (unsigned i = b->nfilters; i < n; i++) b->f[i] = orc_biquad_create(b->nch);
  b->nfilters = n;
}

void orc_fbank_add_filter(orc_fbank* b, const double* c5) {
  orc_fbank_set_filters(b, b->nfilters + 1);
  orc_biquad_set_coeffs(b->f[b->nfilters - 1], c5, 0.0);
}

void orc_fbank_set_channels(orc_fbank* b, unsigned n) {
  for (unsigned i = 0; i < b->nfilters; i++) {
    orc_biquad* q = b->f[i];
    double* w = (double*)calloc(2 * (size_t)(n ? n : 1), sizeof(double));
    memcpy(w, q->w, 2 * (size_t)(n < q->nch ? n : q->nch) * sizeof(double));
    free(q->w);
    q->w = w;
    q->nch = n;
  }
  b->nch = n;
}

void orc_fbank_set_coeffs(orc_fbank* b, unsigned filter, const double* c5, double interp_samples) {
  if (filter < b->nfilters) orc_biquad_set_coeffs(b->f[filter], c5, interp_samples);
}

void orc_fbank_calc(orc_fbank* b, unsigned filter, int type, double freq, double fs, double gain, double bandwidth,
                    double interp_time) {
  if (filter < b->nfilters) orc_biquad_calc(b->f[filter], type, freq, fs, gain, bandwidth, interp_time);
}

void orc_fbank_process(orc_fbank* b, const float* src, float* dst, unsigned nchannels, unsigned nsrc, unsigned ndst,
                       unsigned nframes) {
  if (nchannels > nsrc) nchannels = nsrc;
  if (nchannels > ndst) nchannels = ndst;
  for (unsigned i = 0; i < b->nfilters; i++) {
    orc_biquad_process(b->f[i], src, dst, nchannels, nsrc, ndst, nframes);
    src = dst; /* every later filter works in place on the destination (src/BiQuad.cpp:656-660) */
    nsrc = ndst;
  }
}

void orc_fbank_get_state(const orc_fbank* b, unsigned filter, double* w, double* cur5, double* mul_dec) {
  if (filter < b->nfilters) orc_biquad_get_state(b->f[filter], w, cur5, mul_dec);
}

void orc_fbank_reset(orc_fbank* b) {
  for (unsigned i = 0; i < b->nfilters; i++) orc_biquad_reset(b->f[i]);
}
