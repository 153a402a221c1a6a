/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * cascade.c -- C restatement of BiQuadCascade ("next" row), one cascade per channel of a bank:
 *   construction / numfilters checks   src/BiQuad.h:395-416, :498-511  more than 12 filters -> 0 filters; vectorise needs a
 *                                                                      multiple of four filters, else it is switched off
 *   SetCoefficients (interleaved)      src/BiQuad.h:531-558            (g, b1[0], b2[0], a1[0], a2[0], b1[1], ...), resets the
 *                                                                      registers; g is stored and never applied
 *   one biquad (transposed DF-II)      src/BiQuad.h:667-672            y = x + w0; w0 = x b1 - y a1 + w1; w1 = x b2 - y a2
 *   Tick, cascade form                 src/BiQuad.h:703-715            filter i reads y[i-1] of the same sample
 *   Tick, vectorised form              src/BiQuad.h:687-701, :600-660  every filter reads the x register written by its
 *                                                                      predecessor on the PREVIOUS sample (pipeline: a delay
 *                                                                      of numfilters - 1 samples), then x[1..] = y[0..]
 *   ProcessCascade                     src/BiQuad.h:718-739            Tick per sample (the unrolled form is the same arithmetic)
 * Pinned against the reference's own header compiled (SSE3 intrinsics path) into oracle/_ref (tests/test_cascade.py) and
 * tests/golden/cascade.npz.  float arithmetic, every product, difference and sum rounded separately (-ffp-contract=off).
 * The reference's constructors leave the registers uninitialised; like the tests' reference wrapper this restatement starts
 * from Reset().
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define CASC_MAX 12

typedef struct {
  float b1[CASC_MAX], b2[CASC_MAX], a1[CASC_MAX], a2[CASC_MAX];
  float x[CASC_MAX], y[CASC_MAX], w0[CASC_MAX], w1[CASC_MAX];
  float lastoutput, g;
} casc_one;

struct orc_cascade {
  unsigned nch, nf;
  int vectorise;
  casc_one* c;
};

static void casc_reset(casc_one* c) {
  memset(c->x, 0, sizeof(c->x));
  memset(c->y, 0, sizeof(c->y));
  memset(c->w0, 0, sizeof(c->w0));
  memset(c->w1, 0, sizeof(c->w1));
  c->lastoutput = 0.0f;
}

orc_cascade* orc_cascade_create(unsigned nchannels, unsigned numfilters, int vectorise, int unroll) {
  (void)unroll; /* same arithmetic either way */
  orc_cascade* b = (orc_cascade*)calloc(1, sizeof(*b));
  b->nch = nchannels;
  b->nf = numfilters > CASC_MAX ? 0 : numfilters;            /* src/BiQuad.h:400-404 */
  b->vectorise = (vectorise && (numfilters % 4) == 0) ? 1 : 0; /* src/BiQuad.h:405-409 (tested on the requested count) */
  b->c = (casc_one*)calloc(nchannels ? nchannels : 1, sizeof(casc_one));
  for (unsigned j = 0; j < nchannels; j++) b->c[j].g = 1.0f;
  return b;
}

void orc_cascade_destroy(orc_cascade* b) {
  if (!b) return;
  free(b->c);
  free(b);
}

int orc_cascade_set_coefficients(orc_cascade* b, unsigned channel, const float* coeffs, unsigned n) {
  if (n != 4 * b->nf + 1) return 0; /* src/BiQuad.h:533-537 */
  for (unsigned j = 0; j < b->nch; j++) {
    if (channel != ~0u && channel != j) continue;
    casc_one* c = &b->c[j];
    const float* p = coeffs;
    c->g = *p++;
    for (unsigned i = 0; i < b->nf; i++) {
      c->b1[i] = *p++;
      c->b2[i] = *p++;
      c->a1[i] = *p++;
      c->a2[i] = *p++;
    }
    casc_reset(c);
  }
  return 1;
}

void orc_cascade_reset(orc_cascade* b) {
  for (unsigned j = 0; j < b->nch; j++) casc_reset(&b->c[j]);
}

static float casc_tick(casc_one* c, unsigned nf, int vectorise, float in) {
  if (vectorise) {
    c->x[0] = in;
    for (unsigned i = 0; i + 4 <= nf; i += 4)
      for (unsigned k = i; k < i + 4; k++) {
        const float xv = c->x[k];
        const float yv = xv + c->w0[k];
        c->y[k] = yv;
        c->w0[k] = (xv * c->b1[k] - yv * c->a1[k]) + c->w1[k];
        c->w1[k] = xv * c->b2[k] - yv * c->a2[k];
      }
    memmove(&c->x[1], &c->y[0], sizeof(float) * (nf - 1));
    return c->lastoutput = c->y[nf - 1];
  }
  c->y[0] = in + c->w0[0];
  c->w0[0] = (in * c->b1[0] - c->y[0] * c->a1[0]) + c->w1[0];
  c->w1[0] = in * c->b2[0] - c->y[0] * c->a2[0];
  for (unsigned i = 1; i < nf; i++) {
    c->y[i] = c->y[i - 1] + c->w0[i];
    c->w0[i] = (c->y[i - 1] * c->b1[i] - c->y[i] * c->a1[i]) + c->w1[i];
    c->w1[i] = c->y[i - 1] * c->b2[i] - c->y[i] * c->a2[i];
  }
  return c->lastoutput = c->y[nf - 1];
}

void orc_cascade_process(orc_cascade* b, const float* src, long src_cs, long src_fs, float* dst, long dst_cs, long dst_fs,
                         unsigned nframes) {
  if (!b->nf) return; /* the reference indexes y[-1] with zero filters: undefined, not restated */
  for (unsigned j = 0; j < b->nch; j++)
    for (unsigned i = 0; i < nframes; i++)
      dst[(long)j * dst_cs + (long)i * dst_fs] = casc_tick(&b->c[j], b->nf, b->vectorise, src[(long)j * src_cs + (long)i * src_fs]);
}

unsigned orc_cascade_get_state(const orc_cascade* b, unsigned channel, float* x12, float* y12, float* w0_12, float* w1_12,
                               float* last) {
  if (channel >= b->nch) return 0;
  const casc_one* c = &b->c[channel];
  memcpy(x12, c->x, sizeof(c->x));
  memcpy(y12, c->y, sizeof(c->y));
  memcpy(w0_12, c->w0, sizeof(c->w0));
  memcpy(w1_12, c->w1, sizeof(c->w1));
  *last = c->lastoutput;
  return b->nf | ((unsigned)b->vectorise << 8);
}
