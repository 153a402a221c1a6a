/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Own FFT for the CPU oracle / CPU baseline.  The reference's FFT.{h,cpp},
 * FFT_FFTW.cpp, FFT_kiss.cpp (README:46-51) are absent from the mounted tree and
 * FFTW3 (debian/control:5, libfftw3-dev >= 3.0.0) is not installed, so this file
 * restates the published algorithm FFTW's r2c/c2r interface computes:
 *   forward  X[k] = sum_n x[n] exp(-2 pi i n k / N)          (unnormalised)
 *   inverse  x[n] = sum_k X[k] exp(+2 pi i n k / N)          (unnormalised, caller applies 1/N)
 * fp32 data, twiddles computed in double and rounded once to fp32.
 * Radix-2 decimation-in-time with a bit-reversal table; the real transforms run a
 * half-size complex FFT plus the usual even/odd split.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>

#define MAX_LOG2 20

typedef struct {
  unsigned n;
  unsigned* rev;  /* bit reversal permutation */
  float* tw;      /* n/2 complex twiddles exp(-2 pi i j / n) */
  float* split;   /* n complex split twiddles exp(-2 pi i j / (2n)), j < n  (for the real transform of size 2n) */
} fft_plan;

static fft_plan g_plans[MAX_LOG2 + 1];

static const fft_plan* get_plan(unsigned n) {
  unsigned lg = 0, i;
  while ((1u << lg) < n) lg++;
  fft_plan* p = &g_plans[lg];
  if (p->n == n) return p;
#pragma omp critical(orc_fft_plan)
  {
    if (p->n != n) {
      const double PI = 3.14159265358979323846264338327950288;
      unsigned* rev = (unsigned*)malloc(sizeof(unsigned) * n);
      float* tw = (float*)malloc(sizeof(float) * (n ? n : 1));
      float* split = (float*)malloc(sizeof(float) * 2 * n);
      for (i = 0; i < n; i++) {
        unsigned r = 0, b;
        for (b = 0; b < lg; b++)
          if (i & (1u << b)) r |= 1u << (lg - 1 - b);
        rev[i] = r;
      }
      for (i = 0; i < n / 2; i++) {
        double a = -2.0 * PI * (double)i / (double)n;
        tw[2 * i] = (float)cos(a);
        tw[2 * i + 1] = (float)sin(a);
      }
      for (i = 0; i < n; i++) {
        double a = -PI * (double)i / (double)n;
        split[2 * i] = (float)cos(a);
        split[2 * i + 1] = (float)sin(a);
      }
      p->rev = rev;
      p->tw = tw;
      p->split = split;
#pragma omp flush
      p->n = n;
    }
  }
  return p;
}

void orc_cfft(float* d, unsigned n, int inverse) {
  if (n < 2) return;
  const fft_plan* p = get_plan(n);
  unsigned i, len, j, k;
  for (i = 0; i < n; i++) {
    unsigned r = p->rev[i];
    if (r > i) {
      float tr = d[2 * i], ti = d[2 * i + 1];
      d[2 * i] = d[2 * r];
      d[2 * i + 1] = d[2 * r + 1];
      d[2 * r] = tr;
      d[2 * r + 1] = ti;
    }
  }
  const float sgn = inverse ? -1.0f : 1.0f;
  for (len = 2; len <= n; len <<= 1) {
    unsigned half = len >> 1, step = n / len;
    for (i = 0; i < n; i += len) {
      for (j = 0, k = 0; j < half; j++, k += step) {
        float wr = p->tw[2 * k], wi = sgn * p->tw[2 * k + 1];
        float* a = d + 2 * (i + j);
        float* b = d + 2 * (i + j + half);
        float xr = b[0] * wr - b[1] * wi;
        float xi = b[0] * wi + b[1] * wr;
        b[0] = a[0] - xr;
        b[1] = a[1] - xi;
        a[0] = a[0] + xr;
        a[1] = a[1] + xi;
      }
    }
  }
}

void orc_rfft(const float* in, float* out, unsigned n) {
  unsigned h = n / 2, k;
  const fft_plan* p = get_plan(h);
  /* z[m] = x[2m] + i x[2m+1] is just the input reinterpreted */
  for (k = 0; k < n; k++) out[k] = in[k];
  orc_cfft(out, h, 0);
  float z0r = out[0], z0i = out[1];
  /* pairs (k, h-k) */
  for (k = 1; k <= h / 2; k++) {
    unsigned m = h - k;
    float ar = out[2 * k], ai = out[2 * k + 1];
    float br = out[2 * m], bi = out[2 * m + 1];
    /* E = (A + conj B)/2, O = -i (A - conj B)/2 */
    float er = 0.5f * (ar + br), ei = 0.5f * (ai - bi);
    float dr = 0.5f * (ar - br), di = 0.5f * (ai + bi); /* (A - conj B)/2 */
    float or_ = di, oi = -dr;
    float wr = p->split[2 * k], wi = p->split[2 * k + 1];
    float tr = or_ * wr - oi * wi, ti = or_ * wi + oi * wr;
    out[2 * k] = er + tr;
    out[2 * k + 1] = ei + ti;
    /* X[h-k] = conj(E - w^k O) */
    out[2 * m] = er - tr;
    out[2 * m + 1] = -(ei - ti);
  }
  out[0] = z0r + z0i;
  out[1] = 0.0f;
  out[2 * h] = z0r - z0i;
  out[2 * h + 1] = 0.0f;
}

void orc_irfft(const float* in, float* out, unsigned n) {
  unsigned h = n / 2, k;
  const fft_plan* p = get_plan(h);
  /* Z[k] = (X[k] + conj X[h-k]) + i conj(w^k) (X[k] - conj X[h-k]) */
  for (k = 0; k <= h / 2; k++) {
    unsigned m = h - k;
    float ar = in[2 * k], ai = in[2 * k + 1];
    float br = in[2 * m], bi = in[2 * m + 1];
    if (k == 0) { ai = 0.0f; bi = 0.0f; } /* c2r ignores the imaginary parts of DC and Nyquist */
    float er = ar + br, ei = ai - bi;
    float dr = ar - br, di = ai + bi;
    float wr = p->split[2 * k], wi = -p->split[2 * k + 1]; /* conj(w^k) */
    float tr = dr * wr - di * wi, ti = dr * wi + di * wr;  /* conj(w^k) D */
    /* i * t = (-ti, tr) */
    float zkr = er - ti, zki = ei + tr;
    /* Z[h-k] = conj(E) + i conj(w^(h-k)) (conj of ...) => conj(E - i t)  */
    float zmr = er + ti, zmi = -(ei - tr);
    if (k < h) {
      out[2 * k] = zkr;
      out[2 * k + 1] = zki;
    }
    if (m < h && m != k) {
      out[2 * m] = zmr;
      out[2 * m + 1] = zmi;
    }
  }
  orc_cfft(out, h, 1);
}
