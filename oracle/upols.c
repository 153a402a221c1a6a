/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Uniform-partitioned overlap-save (UPOLS) BlockConvolver, fp32, exactly as
 * SURVEY.md 8.A specifies.  The reference's BlockConvolver.{h,cpp} (README:38-39,
 * "Single-channel partitioned convolution") and simd_utils (README:68-69) are NOT in
 * the mounted tree, so there is no file:line to follow; this is the normative
 * restatement the GPU path is checked against, itself pinned against direct.c
 * (float64 direct convolution) and numpy/scipy float64 in tests/.
 *
 *   B block, N = 2B, K = B+1, P = ceil(L/B)
 *   filter  : H[p][k] = R2C_N([h[pB .. pB+B-1], 0^B])[k]
 *   convolve: FDL[head] = R2C_N([prev, in]); prev = in
 *             Y[k] = sum_{p<P} H[p][k] * FDL[(head - p) mod Pmax][k]     (p ascending)
 *             y = C2R_N(Y) / N ; out = y[B .. 2B-1]
 *             pending filter: with crossfade out = (1-g) o_f + g o_f', g_n = n/B, then f <- f'
 *             head = (head + 1) mod Pmax
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

struct orc_filter {
  unsigned block, partitions;
  float* spectra; /* [P][K] interleaved complex */
};

orc_filter* orc_filter_create(const float* ir, unsigned length, unsigned block) {
  orc_filter* f = (orc_filter*)calloc(1, sizeof(*f));
  unsigned B = block, N = 2 * B, K = B + 1, P = (length + B - 1) / B, p, i;
  if (P == 0) P = 1;
  f->block = B;
  f->partitions = P;
  f->spectra = (float*)calloc((size_t)P * K * 2, sizeof(float));
  float* tmp = (float*)malloc(sizeof(float) * N);
  for (p = 0; p < P; p++) {
    for (i = 0; i < N; i++) {
      size_t idx = (size_t)p * B + i;
      tmp[i] = (i < B && idx < length) ? ir[idx] : 0.0f;
    }
    orc_rfft(tmp, f->spectra + (size_t)p * K * 2, N);
  }
  free(tmp);
  return f;
}

void orc_filter_destroy(orc_filter* f) {
  if (!f) return;
  free(f->spectra);
  free(f);
}
unsigned orc_filter_partitions(const orc_filter* f) { return f->partitions; }
unsigned orc_filter_block(const orc_filter* f) { return f->block; }
const float* orc_filter_spectra(const orc_filter* f) { return f->spectra; }

/* acc[k] += h[k] * x[k], interleaved complex, K bins (the absent simd_utils' job) */
__attribute__((target_clones("avx2", "default")))
void orc_cmac(float* __restrict__ acc, const float* __restrict__ h, const float* __restrict__ x, unsigned K) {
  unsigned k;
  for (k = 0; k < K; k++) {
    float hr = h[2 * k], hi = h[2 * k + 1], xr = x[2 * k], xi = x[2 * k + 1];
    acc[2 * k] += hr * xr - hi * xi;
    acc[2 * k + 1] += hr * xi + hi * xr;
  }
}

struct orc_blockconv {
  unsigned block, pmax, head;
  float* prev;    /* [B] */
  float* fdl;     /* [Pmax][K] complex */
  float* window;  /* [N] */
  float* acc;     /* [K] complex */
  float* y;       /* [N] */
  float* o2;      /* [B] second result on crossfade blocks */
  const orc_filter* cur;
  const orc_filter* pending;
  int has_pending, xfade;
};

orc_blockconv* orc_blockconv_create(unsigned block, unsigned max_partitions) {
  orc_blockconv* bc = (orc_blockconv*)calloc(1, sizeof(*bc));
  unsigned B = block, N = 2 * B, K = B + 1;
  if (max_partitions == 0) max_partitions = 1;
  bc->block = B;
  bc->pmax = max_partitions;
  bc->prev = (float*)calloc(B, sizeof(float));
  bc->fdl = (float*)calloc((size_t)max_partitions * K * 2, sizeof(float));
  bc->window = (float*)calloc(N, sizeof(float));
  bc->acc = (float*)calloc((size_t)K * 2, sizeof(float));
  bc->y = (float*)calloc(N, sizeof(float));
  bc->o2 = (float*)calloc(B, sizeof(float));
  return bc;
}

void orc_blockconv_destroy(orc_blockconv* bc) {
  if (!bc) return;
  free(bc->prev);
  free(bc->fdl);
  free(bc->window);
  free(bc->acc);
  free(bc->y);
  free(bc->o2);
  free(bc);
}

void orc_blockconv_set_filter(orc_blockconv* bc, const orc_filter* f, int crossfade) {
  bc->pending = f;
  bc->has_pending = 1;
  bc->xfade = crossfade;
}

/* one filter against the current FDL -> out[B] */
static void apply_filter(orc_blockconv* bc, const orc_filter* f, float* out) {
  unsigned B = bc->block, N = 2 * B, K = B + 1, p, n;
  if (!f) {
    memset(out, 0, sizeof(float) * B);
    return;
  }
  unsigned P = f->partitions < bc->pmax ? f->partitions : bc->pmax;
  memset(bc->acc, 0, sizeof(float) * 2 * K);
  for (p = 0; p < P; p++) {
    unsigned slot = (bc->head + bc->pmax - p) % bc->pmax;
    orc_cmac(bc->acc, f->spectra + (size_t)p * K * 2, bc->fdl + (size_t)slot * K * 2, K);
  }
  orc_irfft(bc->acc, bc->y, N);
  const float scale = 1.0f / (float)N;
  for (n = 0; n < B; n++) out[n] = bc->y[B + n] * scale;
}

void orc_blockconv_convolve(orc_blockconv* bc, const float* in, float* out) {
  unsigned B = bc->block, N = 2 * B, K = B + 1, n;
  memcpy(bc->window, bc->prev, sizeof(float) * B);
  memcpy(bc->window + B, in, sizeof(float) * B);
  memcpy(bc->prev, in, sizeof(float) * B);
  orc_rfft(bc->window, bc->fdl + (size_t)bc->head * K * 2, N);

  if (bc->has_pending && !bc->xfade) {
    bc->cur = bc->pending;
    bc->has_pending = 0;
  }
  apply_filter(bc, bc->cur, out);
  if (bc->has_pending) { /* crossfaded switch: both filters see the same FDL */
    apply_filter(bc, bc->pending, bc->o2);
    const float inc = 1.0f / (float)B;
    for (n = 0; n < B; n++) {
      float g = (float)n * inc; /* exact for power-of-two B; == Interpolator ramp sampled before the step */
      float a = (1.0f - g) * out[n];
      float b = g * bc->o2[n];
      out[n] = a + b;
    }
    bc->cur = bc->pending;
    bc->has_pending = 0;
  }
  bc->head = (bc->head + 1) % bc->pmax;
  (void)N;
}
