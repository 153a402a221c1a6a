/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * allpass.c -- C restatement of AllPassFilter<float> / AllPassFilterChain<float> ("next" row):
 *   one section      src/AllPassFilter.h:60-68   y[n] = c x[n] + w[n-d];  w[n] = x[n] - c y[n]
 *   block processing src/AllPassFilter.h:84-128  interleaved frames, one ring of nchannels * delay items whose position
 *                                                advances once per channel sample (src/RingBuffer.h:46-53, :88-98);
 *                                                channels that do not fit the src/dst geometry are skipped with Advance
 *   chain            src/AllPassFilter.h:238-255 section 0 reads src, the following sections run in place on dst
 * Pinned against the reference's own headers compiled into oracle/_ref (tests/test_allpass.py) and tests/golden/allpass.npz.
 * float arithmetic, every product and sum rounded separately (-ffp-contract=off).  delay >= 1 and nchannels >= 1 are
 * required (the reference divides by the ring length).
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

struct orc_allpass {
  unsigned nch, nf;
  unsigned* delay;
  float* coeff;
  float** ring; /* [nf][nch * delay] */
  unsigned* pos;
};

orc_allpass* orc_allpass_create(unsigned nchannels, unsigned nfilters, const unsigned* delays, const float* coeffs) {
  orc_allpass* a = (orc_allpass*)calloc(1, sizeof(*a));
  a->nch = nchannels;
  a->nf = nfilters;
  a->delay = (unsigned*)calloc(nfilters ? nfilters : 1, sizeof(unsigned));
  a->coeff = (float*)calloc(nfilters ? nfilters : 1, sizeof(float));
  a->ring = (float**)calloc(nfilters ? nfilters : 1, sizeof(float*));
  a->pos = (unsigned*)calloc(nfilters ? nfilters : 1, sizeof(unsigned));
  for (unsigned f = 0; f < nfilters; f++) {
    a->delay[f] = delays ? delays[f] : 0;
    a->coeff[f] = coeffs ? coeffs[f] : 0.0f;
    a->ring[f] = (float*)calloc((size_t)nchannels * a->delay[f] + 1, sizeof(float));
  }
  return a;
}

void orc_allpass_destroy(orc_allpass* a) {
  if (!a) return;
  for (unsigned f = 0; f < a->nf; f++) free(a->ring[f]);
  free(a->ring);
  free(a->delay);
  free(a->coeff);
  free(a->pos);
  free(a);
}

static void section(orc_allpass* a, unsigned f, const float* src, float* dst, unsigned srcchannel, unsigned nsrc,
                    unsigned dstchannel, unsigned ndst, unsigned nframes) {
  const unsigned len = a->nch * a->delay[f];
  const float c = a->coeff[f];
  float* ring = a->ring[f];
  unsigned pos = a->pos[f];
  if (!len) return;
  src += srcchannel;
  dst += dstchannel;
  unsigned n = a->nch;
  if (a->nch != 1) { /* the single-channel branch of the reference does not clamp */
    const unsigned sa = nsrc >= srcchannel ? nsrc - srcchannel : 0, da = ndst >= dstchannel ? ndst - dstchannel : 0;
    if (n > sa) n = sa;
    if (n > da) n = da;
  }
  for (unsigned i = 0; i < nframes; i++, src += nsrc, dst += ndst) {
    for (unsigned j = 0; j < n; j++) {
      const float x = src[j];
      const float y = c * x + ring[pos];
      dst[j] = y;
      ring[pos] = x - c * y;
      if (++pos >= len) pos = 0;
    }
    pos = (pos + (a->nch - n)) % len;
  }
  a->pos[f] = pos;
}

void orc_allpass_process(orc_allpass* a, const float* src, float* dst, unsigned srcchannel, unsigned nsrc, unsigned dstchannel,
                         unsigned ndst, unsigned nframes) {
  for (unsigned f = 0; f < a->nf; f++) {
    if (f == 1) {
      src = dst;
      srcchannel = dstchannel;
      nsrc = ndst;
    }
    section(a, f, src, dst, srcchannel, nsrc, dstchannel, ndst, nframes);
  }
}

unsigned orc_allpass_get_state(const orc_allpass* a, unsigned f, float* ring, unsigned maxitems) {
  if (f >= a->nf) return 0;
  unsigned n = a->nch * a->delay[f];
  if (n > maxitems) n = maxitems;
  if (n) memcpy(ring, a->ring[f], n * sizeof(float));
  return a->pos[f];
}
