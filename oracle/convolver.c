/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Multichannel Convolver per SURVEY.md 8.A.  The reference's Convolver.{h,cpp}
 * (README:43-44, "Multi-channel parallelized convolution using BlockConvolver", one
 * worker thread per convolver, README:173) is NOT in the mounted tree; this file is
 * the normative CPU restatement.  It composes pieces that ARE pinned to reference code:
 *   I/O           orc_transfer_samples   (SoundFormatConversions.cpp:151-198)
 *   mixdown       orc_mix_samples_f32    (SoundMixing.h:55-81), paths ascending
 *   crossfade     g_n = n/B              (SoundMixing.cpp:43-50 + Interpolator.h:55)
 *   integer delay ring[(w + n - d) mod R](SoundDelayBuffer.cpp:141)
 *   frac. delay   orc_fractional_sample_f32(ring, 0, 1, R, fmod((w + n + R) - d, R))
 *                                        (FractionalSample.cpp:312-341)
 * Per block: inputs -> FDL (one per INPUT, shared by its paths); per path
 * MAC + C2R (+ filter crossfade) -> delay ring -> delayed read (+ delay crossfade)
 * -> MixSamples into the output bus -> TransferSamples out.
 * MIMO mode sums sum_i sum_p H[o][i][p] * FDL_i in the frequency domain (i, then p,
 * ascending) and runs one C2R per output; delays are not available in that mode.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

void orc_cmac(float* acc, const float* h, const float* x, unsigned K); /* upols.c */

typedef struct {
  unsigned input, output;
  float gain;
  const orc_filter* cur;
  const orc_filter* pend;
  int has_pending, xfade;
  double delay, pend_delay;
  float* ring;
} path_t;

struct orc_convolver {
  unsigned B, pmax, n_in, n_out, n_paths, R, head, w;
  int mode, fractional, nthreads;
  float* prev;   /* [n_in][B] */
  float* fdl;    /* [n_in][Pmax][K] complex */
  path_t* paths;
  float* xin;    /* [n_in][B]   de-interleaved input block */
  float* pout;   /* [n_paths|n_out][B] per-path output after delay */
  float* bus;    /* [B][n_out] */
  float* scratch; /* per thread: window[N] + acc[2K] + y[N] + o1[B] + o2[B] */
  size_t scratch_stride;
};

orc_convolver* orc_convolver_create(unsigned block, unsigned max_partitions, unsigned n_inputs, unsigned n_outputs,
                                    unsigned n_paths, int mode, unsigned ring_len, int fractional_delay,
                                    int nthreads) {
  orc_convolver* c = (orc_convolver*)calloc(1, sizeof(*c));
  unsigned B = block, N = 2 * B, K = B + 1, i;
  if (mode == ORC_MODE_PER_CHANNEL) { n_outputs = n_inputs; n_paths = n_inputs; }
  if (mode == ORC_MODE_MIMO) n_paths = n_inputs * n_outputs;
  if (max_partitions == 0) max_partitions = 1;
  if (ring_len < B) ring_len = B;
  if (nthreads < 1) nthreads = 1;
  c->B = B; c->pmax = max_partitions; c->n_in = n_inputs; c->n_out = n_outputs; c->n_paths = n_paths;
  c->R = ring_len; c->mode = mode; c->fractional = fractional_delay; c->nthreads = nthreads;
  c->prev = (float*)calloc((size_t)n_inputs * B, sizeof(float));
  c->fdl = (float*)calloc((size_t)n_inputs * max_partitions * K * 2, sizeof(float));
  c->paths = (path_t*)calloc(n_paths, sizeof(path_t));
  unsigned nrings = (mode == ORC_MODE_MIMO) ? 0 : n_paths;
  for (i = 0; i < n_paths; i++) {
    path_t* p = &c->paths[i];
    if (mode == ORC_MODE_MIMO) { p->output = i / n_inputs; p->input = i % n_inputs; }
    else if (mode == ORC_MODE_PER_CHANNEL) { p->input = i; p->output = i; }
    p->gain = 1.0f;
    if (i < nrings) p->ring = (float*)calloc(ring_len, sizeof(float));
  }
  c->xin = (float*)calloc((size_t)n_inputs * B, sizeof(float));
  unsigned npo = (mode == ORC_MODE_MIMO) ? n_outputs : n_paths;
  c->pout = (float*)calloc((size_t)npo * B, sizeof(float));
  c->bus = (float*)calloc((size_t)B * n_outputs, sizeof(float));
  c->scratch_stride = (size_t)N + 2 * K + N + B + B + 16;
  c->scratch = (float*)calloc(c->scratch_stride * (size_t)nthreads, sizeof(float));
  return c;
}

void orc_convolver_destroy(orc_convolver* c) {
  unsigned i;
  if (!c) return;
  for (i = 0; i < c->n_paths; i++) free(c->paths[i].ring);
  free(c->prev); free(c->fdl); free(c->paths); free(c->xin); free(c->pout); free(c->bus); free(c->scratch);
  free(c);
}

void orc_convolver_set_route(orc_convolver* c, unsigned path, unsigned input, unsigned output, float gain) {
  if (path >= c->n_paths || c->mode != ORC_MODE_ROUTED) return;
  if (input >= c->n_in || output >= c->n_out) return;
  c->paths[path].input = input;
  c->paths[path].output = output;
  c->paths[path].gain = gain;
}

void orc_convolver_set_filter(orc_convolver* c, unsigned path, const orc_filter* f, int crossfade, double delay) {
  if (path >= c->n_paths) return;
  path_t* p = &c->paths[path];
  p->pend = f;
  p->pend_delay = delay;
  p->xfade = crossfade;
  p->has_pending = 1;
}

#ifdef _OPENMP
#include <omp.h>
static int thread_id(void) { return omp_get_thread_num(); }
#else
static int thread_id(void) { return 0; }
#endif

/* acc[K] (+)= sum_p H_f[p] * FDL_input[(head - p) mod Pmax] */
static void mac_filter(const orc_convolver* c, const orc_filter* f, unsigned input, float* acc) {
  unsigned K = c->B + 1, p;
  if (!f) return;
  unsigned P = orc_filter_partitions(f) < c->pmax ? orc_filter_partitions(f) : c->pmax;
  const float* H = orc_filter_spectra(f);
  const float* fdl = c->fdl + (size_t)input * c->pmax * K * 2;
  for (p = 0; p < P; p++) {
    unsigned slot = (c->head + c->pmax - p) % c->pmax;
    orc_cmac(acc, H + (size_t)p * K * 2, fdl + (size_t)slot * K * 2, K);
  }
}

static void acc_to_block(const orc_convolver* c, float* acc, float* y, float* out) {
  unsigned B = c->B, N = 2 * B, n;
  const float scale = 1.0f / (float)N;
  orc_irfft(acc, y, N);
  for (n = 0; n < B; n++) out[n] = y[B + n] * scale; /* overlap-save: discard the first B samples */
}

static void crossfade_block(unsigned B, float* o1, const float* o2) {
  unsigned n;
  const float inc = 1.0f / (float)B;
  for (n = 0; n < B; n++) {
    float g = (float)n * inc;
    float a = (1.0f - g) * o1[n];
    float b = g * o2[n];
    o1[n] = a + b;
  }
}

static float delayed_read(const orc_convolver* c, const float* ring, unsigned n, double d) {
  unsigned R = c->R;
  if (c->fractional) {
    double pos = fmod((double)(c->w + n + R) - d, (double)R);
    return (float)orc_fractional_sample_f32(ring, 0, 1, R, pos);
  }
  unsigned di = (unsigned)d;
  return ring[(c->w + n + R - (di % R)) % R];
}

static void process_block(orc_convolver* c) {
  const unsigned B = c->B, N = 2 * B, K = B + 1;
  int i;

  /* 1. inputs -> FDL[head] */
#pragma omp parallel for num_threads(c->nthreads) schedule(static)
  for (i = 0; i < (int)c->n_in; i++) {
    float* s = c->scratch + c->scratch_stride * (size_t)thread_id();
    float* window = s;
    float* prev = c->prev + (size_t)i * B;
    const float* cur = c->xin + (size_t)i * B;
    memcpy(window, prev, sizeof(float) * B);
    memcpy(window + B, cur, sizeof(float) * B);
    memcpy(prev, cur, sizeof(float) * B);
    orc_rfft(window, c->fdl + ((size_t)i * c->pmax + c->head) * K * 2, N);
  }

  memset(c->bus, 0, sizeof(float) * (size_t)B * c->n_out);

  if (c->mode == ORC_MODE_MIMO) {
#pragma omp parallel for num_threads(c->nthreads) schedule(dynamic, 1)
    for (i = 0; i < (int)c->n_out; i++) {
      float* s = c->scratch + c->scratch_stride * (size_t)thread_id();
      float* acc = s + N;
      float* y = acc + 2 * K;
      float* o1 = c->pout + (size_t)i * B;
      float* o2 = y + N + B;
      unsigned in, any_xfade = 0;
      /* non-crossfaded switches apply first */
      for (in = 0; in < c->n_in; in++) {
        path_t* p = &c->paths[(size_t)i * c->n_in + in];
        if (p->has_pending && !p->xfade) { p->cur = p->pend; p->has_pending = 0; }
        if (p->has_pending) any_xfade = 1;
      }
      memset(acc, 0, sizeof(float) * 2 * K);
      for (in = 0; in < c->n_in; in++) mac_filter(c, c->paths[(size_t)i * c->n_in + in].cur, in, acc);
      acc_to_block(c, acc, y, o1);
      if (any_xfade) {
        memset(acc, 0, sizeof(float) * 2 * K);
        for (in = 0; in < c->n_in; in++) {
          path_t* p = &c->paths[(size_t)i * c->n_in + in];
          mac_filter(c, p->has_pending ? p->pend : p->cur, in, acc);
        }
        acc_to_block(c, acc, y, o2);
        crossfade_block(B, o1, o2);
        for (in = 0; in < c->n_in; in++) {
          path_t* p = &c->paths[(size_t)i * c->n_in + in];
          if (p->has_pending) { p->cur = p->pend; p->has_pending = 0; }
        }
      }
    }
    for (i = 0; i < (int)c->n_out; i++)
      orc_mix_samples_f32(c->pout + (size_t)i * B, 0, 1, c->bus, (unsigned)i, c->n_out, 1, B, 1.0f);
    return;
  }

#pragma omp parallel for num_threads(c->nthreads) schedule(dynamic, 1)
  for (i = 0; i < (int)c->n_paths; i++) {
    float* s = c->scratch + c->scratch_stride * (size_t)thread_id();
    float* acc = s + N;
    float* y = acc + 2 * K;
    float* o1 = y + N;
    float* o2 = o1 + B;
    path_t* p = &c->paths[i];
    float* po = c->pout + (size_t)i * B;
    unsigned n;
    double d_old = p->delay, d_new = p->delay;
    int delay_xfade = 0;

    if (p->has_pending && !p->xfade) { /* hard switch: filter and delay jump at this boundary */
      p->cur = p->pend;
      p->delay = p->pend_delay;
      d_old = d_new = p->delay;
      p->has_pending = 0;
    }
    memset(acc, 0, sizeof(float) * 2 * K);
    mac_filter(c, p->cur, p->input, acc);
    if (p->cur) acc_to_block(c, acc, y, o1); else memset(o1, 0, sizeof(float) * B);
    if (p->has_pending) { /* crossfaded switch */
      memset(acc, 0, sizeof(float) * 2 * K);
      mac_filter(c, p->pend, p->input, acc);
      if (p->pend) acc_to_block(c, acc, y, o2); else memset(o2, 0, sizeof(float) * B);
      crossfade_block(B, o1, o2);
      d_new = p->pend_delay;
      delay_xfade = (d_new != d_old);
      p->cur = p->pend;
      p->delay = p->pend_delay;
      p->has_pending = 0;
    }
    /* delay ring: write the block at w .. w+B-1, then read it back delayed */
    for (n = 0; n < B; n++) p->ring[(c->w + n) % c->R] = o1[n];
    if (delay_xfade) {
      const float inc = 1.0f / (float)B;
      for (n = 0; n < B; n++) {
        float g = (float)n * inc;
        float a = (1.0f - g) * delayed_read(c, p->ring, n, d_old);
        float b = g * delayed_read(c, p->ring, n, d_new);
        po[n] = a + b;
      }
    } else {
      for (n = 0; n < B; n++) po[n] = delayed_read(c, p->ring, n, d_new);
    }
  }
  /* mixdown, paths ascending (deterministic order) */
  for (i = 0; i < (int)c->n_paths; i++) {
    const path_t* p = &c->paths[i];
    orc_mix_samples_f32(c->pout + (size_t)i * B, 0, 1, c->bus, p->output, c->n_out, 1, B, p->gain);
  }
}

int orc_convolver_process(orc_convolver* c, const void* in, int infmt, int in_be, unsigned in_channels, void* out,
                          int outfmt, int out_be, unsigned out_channels, unsigned nframes) {
  unsigned B = c->B, t, i, T = nframes / B;
  if (nframes % B) return -1;
  if (in_channels < c->n_in || out_channels < c->n_out) return -2;
  unsigned inlen = orc_get_bytes_per_sample(infmt), outlen = orc_get_bytes_per_sample(outfmt);
  for (t = 0; t < T; t++) {
    const uint8_t* src = (const uint8_t*)in + (size_t)t * B * in_channels * inlen;
    uint8_t* dst = (uint8_t*)out + (size_t)t * B * out_channels * outlen;
    for (i = 0; i < c->n_in; i++)
      orc_transfer_samples(src, infmt, in_be, i, in_channels, c->xin + (size_t)i * B, ORC_FMT_FLOAT, 0, 0, 1, 1, B);
    process_block(c);
    orc_transfer_samples(c->bus, ORC_FMT_FLOAT, 0, 0, c->n_out, dst, outfmt, out_be, 0, out_channels, c->n_out, B);
    c->head = (c->head + 1) % c->pmax;
    c->w = (c->w + B) % c->R;
  }
  return 0;
}
