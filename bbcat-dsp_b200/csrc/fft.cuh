// fft.cuh -- shared-memory complex FFT of M = B points and the real-transform split stages.
//
// Replaces the reference's FFT abstraction (FFT.{h,cpp}, FFT_FFTW.cpp, FFT_kiss.cpp, README:46-51;
// absent from the mounted tree).  Conventions are FFTW's r2c/c2r: both directions unnormalised,
// forward kernel exp(-2 pi i nk/N); the single 1/N is folded into the filter spectra (exact, N is a
// power of two).
//
// A real transform of N = 2B points is one complex FFT of M = B points on z[n] = x[2n] + i x[2n+1]
// plus an even/odd split.  The complex FFT is a Stockham autosort (natural order in and out, no bit
// reversal) with the data in shared memory between passes and R values per thread in registers:
//   M = 64, 512, 4096 (powers of 8): radix-8 passes, M/8 threads per transform
//   other powers of two            : radix-4 passes (+ one radix-2 pass), M/4 threads per transform
// Several transforms share a CTA (threadIdx.y) so that small sizes still launch 128-256 threads.
// Twiddles come from a table computed in double precision on the host
// (tw[j] = exp(-2 pi i j / N), j < N; the M-point twiddle exp(-2 pi i j / M) is tw[2j]).
// Shared-memory indices go through PADM<M>: one element of padding per eight for the radix-4 sizes, an XOR swizzle
// for the radix-8 sizes, so that the stride-R stores of the first passes are bank-conflict free.
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
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 d) {
  return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
}

template <int M>
struct FftCfg {
  static constexpr int R = (M == 64 || M == 512 || M == 4096) ? 8 : 4;  // values per thread
  static constexpr int NT = M / R;                                     // threads per transform
  static constexpr int FPB = (NT >= 256) ? 1 : ((256 / NT) > 16 ? 16 : (256 / NT));  // transforms per CTA
  static constexpr int MP = M + M / 8;                                  // padded smem elements per transform
};
// Index of element i inside a transform's shared-memory workspace.
//  * radix-4 sizes: one float2 of padding per 8 elements.
//  * radix-8 sizes: an XOR swizzle of the low four index bits with bits 3..6 (a bijection inside every aligned group of
//    128 elements, so the workspace needs no padding; the allocation keeps MP elements per transform so that the
//    transforms of a CTA stay staggered).  A half-warp moves 16 float2 per wavefront, i.e. the bank pair is the low four
//    bits: 16 consecutive elements (every pass's reads, the last passes' writes) keep distinct low bits under the XOR
//    with a constant; the first pass's transposed writes 8 j + r and the second pass's two runs of eight 64 elements apart
//    get theirs from bits 3..6.  With the padding those sequential reads straddled a pad and took two wavefronts each
//    (ncu: 42 % of k_rfft8's shared-memory wavefronts were conflict replays; k_rfft8 19.9 -> 17.1 us per C3 step).
__device__ __forceinline__ int PAD(int i) { return i + (i >> 3); }
template <int M>
__device__ __forceinline__ int PADM(int i) {
  if constexpr (M == 64 || M == 512 || M == 4096) return i ^ ((i >> 3) & 15);
  else return PAD(i);
}

// 4-point DFT, outputs in natural order
template <bool INV>
__device__ __forceinline__ void fft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_mi<INV>(csub(v1, v3));
  v0 = cadd(a0, a2);
  v1 = cadd(a1, a3);
  v2 = csub(a0, a2);
  v3 = csub(a1, a3);
}

// 8-point DFT (decimation in frequency: even outputs from the sums, odd outputs from the twiddled
// differences), outputs in natural order
template <bool INV>
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
  constexpr float S = 0.70710678118654752440f;
  float2 b0 = cadd(v[0], v[4]), b1 = cadd(v[1], v[5]), b2 = cadd(v[2], v[6]), b3 = cadd(v[3], v[7]);
  float2 c0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
  // c_n = d_n * W8^n, W8 = exp(-+ i pi/4): W8^1 = (1 -+ i)/sqrt2, W8^2 = -+ i, W8^3 = (-1 -+ i)/sqrt2
  float2 c1 = INV ? make_float2(S * (d1.x - d1.y), S * (d1.x + d1.y)) : make_float2(S * (d1.x + d1.y), S * (d1.y - d1.x));
  float2 c2 = mul_mi<INV>(d2);
  float2 c3 = INV ? make_float2(-S * (d3.x + d3.y), S * (d3.x - d3.y)) : make_float2(S * (d3.y - d3.x), -S * (d3.x + d3.y));
  fft4<INV>(b0, b1, b2, b3);
  fft4<INV>(c0, c1, c2, c3);
  v[0] = b0;
  v[1] = c0;
  v[2] = b1;
  v[3] = c1;
  v[4] = b2;
  v[5] = c2;
  v[6] = b3;
  v[7] = c3;
}

// Complex FFT of M points held in smem `s` (padded, FftCfg<M>::MP float2); threads tid = 0..NT-1 of one
// transform take part, every thread of the CTA must call it (block-wide barriers).  Result in `s`,
// natural order.  Ends with a __syncthreads().
template <int M, bool INV>
__device__ __forceinline__ void cfft_smem(float2* __restrict__ s, const float2* __restrict__ tw, int tid) {
  constexpr int R = FftCfg<M>::R, NT = FftCfg<M>::NT;
  const int j = tid;
  int Ns = 1;
  if (R == 8) {
#pragma unroll 1
    for (; Ns * 8 <= M; Ns *= 8) {
      const int k = j & (Ns - 1);
      float2 v[8];
#pragma unroll
      for (int r = 0; r < 8; r++) v[r] = s[PADM<M>(j + r * NT)];
      if (Ns > 1) {
        const int stride = 2 * (M / (8 * Ns));  // tw index of exp(-2 pi i k / (8 Ns))
#pragma unroll
        for (int r = 1; r < 8; r++) {
          float2 w = __ldg(&tw[r * k * stride]);
          if (INV) w.y = -w.y;
          v[r] = cmul(v[r], w);
        }
      }
      fft8<INV>(v);
      const int j0 = ((j - k) << 3) + k;
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 8; r++) s[PADM<M>(j0 + r * Ns)] = v[r];
      __syncthreads();
    }
  } else {
#pragma unroll 1
    for (; Ns * 4 <= M; Ns *= 4) {
      const int k = j & (Ns - 1);
      float2 v0 = s[PADM<M>(j)], v1 = s[PADM<M>(j + NT)], v2 = s[PADM<M>(j + 2 * NT)], v3 = s[PADM<M>(j + 3 * NT)];
      if (Ns > 1) {
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
      fft4<INV>(v0, v1, v2, v3);
      const int j0 = ((j - k) << 2) + k;
      __syncthreads();
      s[PADM<M>(j0)] = v0;
      s[PADM<M>(j0 + Ns)] = v1;
      s[PADM<M>(j0 + 2 * Ns)] = v2;
      s[PADM<M>(j0 + 3 * Ns)] = v3;
      __syncthreads();
    }
    if (Ns < M) {
      // one radix-2 pass left (Ns == M/2): each thread does two butterflies
      float2 r[4];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int jj = j + h * NT;  // jj < M/2
        const int k = jj & (Ns - 1);
        float2 v0 = s[PADM<M>(jj)], v1 = s[PADM<M>(jj + M / 2)];
        float2 w = __ldg(&tw[2 * k * (M / (2 * Ns))]);
        if (INV) w.y = -w.y;
        v1 = cmul(v1, w);
        r[2 * h] = cadd(v0, v1);
        r[2 * h + 1] = csub(v0, v1);
      }
      __syncthreads();
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int jj = j + h * NT;
        const int k = jj & (Ns - 1);
        const int j0 = ((jj - k) << 1) + k;
        s[PADM<M>(j0)] = r[2 * h];
        s[PADM<M>(j0 + Ns)] = r[2 * h + 1];
      }
      __syncthreads();
    }
  }
}

// ---- radix-8 sizes (M = 64, 512, 4096): persistent-kernel building blocks ----
// A thread's twiddles depend on its index only, not on the transform, so a CTA that loops over many transforms
// loads them once into registers: tw8.w[p][r-1] = exp(-2 pi i r k / (8 Ns)) for pass p+1 (Ns = 8^(p+1),
// k = j mod Ns) and tw8.ws[h] = exp(-2 pi i (j + h NT) / 2M) for the real-transform split stage.
template <int M>
struct Tw8 {
  static constexpr int NP = (M == 64) ? 1 : (M == 512) ? 2 : 3;  // twiddled passes (the first pass has none)
  float2 w[NP][7];
  float2 ws[8];
};

template <int M>
__device__ __forceinline__ void load_tw8(Tw8<M>& t, const float2* __restrict__ tw, int j) {
  constexpr int NT = M / 8;
  int Ns = 8;
#pragma unroll
  for (int p = 0; p < Tw8<M>::NP; p++) {
    const int k = j & (Ns - 1), stride = 2 * (M / (8 * Ns));
#pragma unroll
    for (int r = 1; r < 8; r++) t.w[p][r - 1] = __ldg(&tw[r * k * stride]);
    Ns *= 8;
  }
#pragma unroll
  for (int h = 0; h < 8; h++) t.ws[h] = __ldg(&tw[j + h * NT]);
}

// barrier over the NT threads of one transform (threadIdx.y): a named barrier when they are whole warps and the CTA
// holds several transforms, otherwise the block barrier (every transform of the CTA runs the same sequence)
template <int M>
__device__ __forceinline__ void fft_bar() {
  constexpr int NT = M / 8;
  if (NT % 32 == 0 && FftCfg<M>::FPB > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)threadIdx.y), "n"(NT) : "memory");
  else __syncthreads();
}

// first pass (Ns = 1, no twiddles) on values already in registers (v[r] = element j + r NT), result to smem
template <int M, bool INV>
__device__ __forceinline__ void pass8_first(float2 (&v)[8], float2* __restrict__ s, int j) {
  fft8<INV>(v);
#pragma unroll
  for (int r = 0; r < 8; r++) s[PADM<M>(8 * j + r)] = v[r];
  fft_bar<M>();
}

// pass with Ns = NS > 1 (smem -> smem), ends with a barrier
template <int M, bool INV, int NS>
__device__ __forceinline__ void pass8(float2* __restrict__ s, const float2 (&w)[7], int j) {
  constexpr int NT = M / 8;
  const int k = j & (NS - 1);
  float2 v[8];
#pragma unroll
  for (int r = 0; r < 8; r++) v[r] = s[PADM<M>(j + r * NT)];
#pragma unroll
  for (int r = 1; r < 8; r++) {
    float2 ww = w[r - 1];
    if (INV) ww.y = -ww.y;
    v[r] = cmul(v[r], ww);
  }
  fft8<INV>(v);
  const int j0 = ((j - k) << 3) + k;
  fft_bar<M>();
#pragma unroll
  for (int r = 0; r < 8; r++) s[PADM<M>(j0 + r * NS)] = v[r];
  fft_bar<M>();
}

// all passes after the first
template <int M, bool INV>
__device__ __forceinline__ void passes8_rest(float2* __restrict__ s, const Tw8<M>& t, int j) {
  pass8<M, INV, 8>(s, t.w[0], j);
  if constexpr (M >= 512) pass8<M, INV, 64>(s, t.w[1], j);
  if constexpr (M >= 4096) pass8<M, INV, 512>(s, t.w[2], j);
}

// forward split with cached twiddles (same arithmetic as rfft_split_store)
template <int M>
__device__ __forceinline__ void rfft_split_store8(const float2* __restrict__ s, const Tw8<M>& t, float2* __restrict__ out,
                                                  float scale, int tid, bool active) {
  constexpr int NT = M / 8;
#pragma unroll
  for (int h = 0; h < 8; h++) {
    const int k = tid + h * NT;
    float2 x;
    if (k == 0) {
      float2 z0 = s[0];
      x = make_float2(z0.x + z0.y, z0.x - z0.y);  // (DC, Nyquist)
    } else {
      float2 a = s[PADM<M>(k)], b = s[PADM<M>(M - k)];
      float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
      float2 o = make_float2(d.y, -d.x);
      x = cadd(e, cmul(o, t.ws[h]));
    }
    if (active) out[k] = make_float2(x.x * scale, x.y * scale);
  }
}

// inverse split into registers: v[h] = Z[tid + h NT] (same arithmetic as irfft_unsplit), x = summed packed spectrum
template <int M>
__device__ __forceinline__ void irfft_unsplit8(const float2* __restrict__ x, const Tw8<M>& t, float2 (&v)[8], int tid) {
  constexpr int NT = M / 8;
#pragma unroll
  for (int h = 0; h < 8; h++) {
    const int k = tid + h * NT;
    float2 z;
    if (k == 0) {
      float2 p = x[0];
      z = make_float2(p.x + p.y, p.x - p.y);
    } else {
      float2 a = x[k], b = x[M - k];
      float2 e = make_float2(a.x + b.x, a.y - b.y);
      float2 d = make_float2(a.x - b.x, a.y + b.y);
      float2 w = t.ws[h];
      w.y = -w.y;
      float2 tt = cmul(d, w);
      z = make_float2(e.x - tt.y, e.y + tt.x);
    }
    v[h] = z;
  }
}

// Forward split: Z = FFT_M(z) in smem -> packed half spectrum X (M complex) of the 2M-point real
// transform, scaled by `scale`, written to global `out` (coalesced float2).
//   X[k] = E + w^k O,  E = (Z[k] + conj Z[M-k])/2,  O = -i (Z[k] - conj Z[M-k])/2,  w = exp(-2 pi i / 2M)
template <int M>
__device__ __forceinline__ void rfft_split_store(const float2* __restrict__ s, const float2* __restrict__ tw,
                                                 float2* __restrict__ out, float scale, int tid, bool active) {
  constexpr int R = FftCfg<M>::R, NT = FftCfg<M>::NT;
#pragma unroll
  for (int h = 0; h < R; h++) {
    const int k = tid + h * NT;  // 0 .. M-1
    float2 x;
    if (k == 0) {
      float2 z0 = s[0];
      x = make_float2(z0.x + z0.y, z0.x - z0.y);  // (DC, Nyquist)
    } else {
      float2 a = s[PADM<M>(k)], b = s[PADM<M>(M - k)];
      float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
      float2 o = make_float2(d.y, -d.x);
      float2 t = cmul(o, __ldg(&tw[k]));
      x = cadd(e, t);
    }
    if (active) out[k] = make_float2(x.x * scale, x.y * scale);
  }
}

// Inverse split: packed half spectrum X (M complex, already summed, unpadded smem `x`) -> Z in padded
// smem `s` such that IFFT_M(Z) = (y[2n] + i y[2n+1]) of an unnormalised c2r.
//   Z[k] = (X[k] + conj X[M-k]) + i conj(w^k) (X[k] - conj X[M-k])
template <int M>
__device__ __forceinline__ void irfft_unsplit(const float2* __restrict__ x, const float2* __restrict__ tw,
                                              float2* __restrict__ s, int tid) {
  constexpr int R = FftCfg<M>::R, NT = FftCfg<M>::NT;
#pragma unroll
  for (int h = 0; h < R; h++) {
    const int k = tid + h * NT;
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
    s[PADM<M>(k)] = z;
  }
}

}  // namespace bbx
