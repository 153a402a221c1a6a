// fft.cuh -- shared-memory complex FFT of M = B points and the real-transform split stages.
//
// Replaces the reference's FFT abstraction (FFT.{h,cpp}, FFT_FFTW.cpp, FFT_kiss.cpp, README:46-51;
// absent from the mounted tree).  Conventions are FFTW's r2c/c2r: both directions unnormalised,
// forward kernel exp(-2 pi i nk/N); the single 1/N is folded into the filter spectra (exact, N is a
// power of two).
//
// A real transform of N = 2B points is one complex FFT of M = B points on z[n] = x[2n] + i x[2n+1]
// plus an even/odd split.  The complex FFT is a Stockham autosort (natural order in and out, no bit
// reversal), radix-4 passes with one radix-2 pass when log2(M) is odd, M/4 threads, data in shared
// memory, twiddles read from a table computed in double precision on the host
// (tw[j] = exp(-2 pi i j / N), j < N; the M-point twiddle exp(-2 pi i j / M) is tw[2j]).
//
// Spectra are stored PACKED: B complex values per row, bin 0 holds (X[0].re, X[B].re) -- DC and
// Nyquist are both real for real input -- so a row is exactly 8B bytes (4 KB at B = 512).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bbx {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// Complex FFT of M points held in smem `s` (float2[M]); blockDim.x == M/4 threads take part.
// INV = false: kernel exp(-i...), INV = true: exp(+i...).  Result in `s`, natural order.
// Ends with a __syncthreads().
template <int M, bool INV>
__device__ __forceinline__ void cfft_smem(float2* __restrict__ s, const float2* __restrict__ tw) {
  constexpr int Q = M / 4;
  const int j = threadIdx.x;
  int Ns = 1;
#pragma unroll 1
  for (; Ns * 4 <= M; Ns *= 4) {
    const int k = j & (Ns - 1);
    float2 v0 = s[j], v1 = s[j + Q], v2 = s[j + 2 * Q], v3 = s[j + 3 * Q];
    if (Ns > 1) {
      // exp(-2 pi i r k / (4 Ns)) = tw_M[r k M/(4 Ns)] = tw[2 r k M/(4 Ns)]
      const int stride = 2 * (M / (4 * Ns));
      float2 w1 = __ldg(&tw[k * stride]), w2 = __ldg(&tw[2 * k * stride]), w3 = __ldg(&tw[3 * k * stride]);
      if (INV) {
        w1.y = -w1.y;
        w2.y = -w2.y;
        w3.y = -w3.y;
      }
      v1 = cmul(v1, w1);
      v2 = cmul(v2, w2);
      v3 = cmul(v3, w3);
    }
    float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), d = csub(v1, v3);
    // forward: (v1 - v3) * (-i) = (d.y, -d.x); inverse: * (+i) = (-d.y, d.x)
    float2 a3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    const int j0 = ((j - k) << 2) + k;
    __syncthreads();
    s[j0] = cadd(a0, a2);
    s[j0 + Ns] = cadd(a1, a3);
    s[j0 + 2 * Ns] = csub(a0, a2);
    s[j0 + 3 * Ns] = csub(a1, a3);
    __syncthreads();
  }
  if (Ns < M) {
    // one radix-2 pass left (Ns == M/2): each thread does two butterflies
    float2 r[4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int jj = j + h * Q;  // jj < M/2
      const int k = jj & (Ns - 1);
      float2 v0 = s[jj], v1 = s[jj + M / 2];
      float2 w = __ldg(&tw[2 * k * (M / (2 * Ns))]);
      if (INV) w.y = -w.y;
      v1 = cmul(v1, w);
      r[2 * h] = cadd(v0, v1);
      r[2 * h + 1] = csub(v0, v1);
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int jj = j + h * Q;
      const int k = jj & (Ns - 1);
      const int j0 = ((jj - k) << 1) + k;
      s[j0] = r[2 * h];
      s[j0 + Ns] = r[2 * h + 1];
    }
    __syncthreads();
  }
}

// Forward split: Z = FFT_M(z) in smem -> packed half spectrum X (M complex) of the 2M-point real
// transform, scaled by `scale`, written to global `out` (coalesced float2).
//   X[k] = E + w^k O,  E = (Z[k] + conj Z[M-k])/2,  O = -i (Z[k] - conj Z[M-k])/2,  w = exp(-2 pi i / 2M)
template <int M>
__device__ __forceinline__ void rfft_split_store(const float2* __restrict__ s, const float2* __restrict__ tw,
                                                 float2* __restrict__ out, float scale) {
  constexpr int Q = M / 4;
  const int j = threadIdx.x;
#pragma unroll
  for (int h = 0; h < 4; h++) {
    const int k = j + h * Q;  // 0 .. M-1
    float2 x;
    if (k == 0) {
      float2 z0 = s[0];
      x = make_float2(z0.x + z0.y, z0.x - z0.y);  // (DC, Nyquist)
    } else {
      float2 a = s[k], b = s[M - k];
      float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
      float2 o = make_float2(d.y, -d.x);
      float2 t = cmul(o, __ldg(&tw[k]));
      x = cadd(e, t);
    }
    out[k] = make_float2(x.x * scale, x.y * scale);
  }
}

// Inverse split: packed half spectrum X (M complex, already summed) -> Z in smem such that
// IFFT_M(Z) = (y[2n] + i y[2n+1]) * 2M-scaling of an unnormalised c2r.
//   Z[k] = (X[k] + conj X[M-k]) + i conj(w^k) (X[k] - conj X[M-k])
// `x` is a smem copy of the packed spectrum; result written to `s` (distinct array).
template <int M>
__device__ __forceinline__ void irfft_unsplit(const float2* __restrict__ x, const float2* __restrict__ tw,
                                              float2* __restrict__ s) {
  constexpr int Q = M / 4;
  const int j = threadIdx.x;
#pragma unroll
  for (int h = 0; h < 4; h++) {
    const int k = j + h * Q;
    float2 z;
    if (k == 0) {
      float2 p = x[0];  // (DC, Nyquist)
      z = make_float2(p.x + p.y, p.x - p.y);
    } else {
      float2 a = x[k], b = x[M - k];
      float2 e = make_float2(a.x + b.x, a.y - b.y);
      float2 d = make_float2(a.x - b.x, a.y + b.y);
      float2 w = __ldg(&tw[k]);
      w.y = -w.y;  // conj(w^k)
      float2 t = cmul(d, w);
      z = make_float2(e.x - t.y, e.y + t.x);
    }
    s[k] = z;
  }
}

}  // namespace bbx
