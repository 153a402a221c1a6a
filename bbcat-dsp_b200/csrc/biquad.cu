// biquad.cu -- BiQuadCoeffs / BiQuad on the GPU (the "next" row after the convolver in a renderer
// chain: EQ banks).  Replaces, for one coefficient object shared by a bank of per-channel filters,
//   BiQuadCoeffs::SetCoeffs / CalcCoeffs / Interpolate      src/BiQuad.cpp:75-103, :181-352, :379-395
//   BiQuad::Process(x)                                       src/BiQuad.h:200-206  (direct form II transposed, double state)
//   BiQuad::Process(filters, src, dst, ...) with the ramp    src/BiQuad.cpp:473-497 (what BiQuadFilterBank::Process runs)
// The recurrence is serial in time, so the parallel axis is the channel: one thread per channel walks the frames,
// every thread advances its own copy of the coefficient ramp (same IEEE operations, same order as the reference: the
// results are bit-exact, products and sums rounded separately, no FMA contraction).  Coefficient design is host
// arithmetic (libm), the state lives in HBM.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace bbx {

struct BiquadCoeffState {
  double cur[5], tgt[5], dif[5];  // num0 num1 num2 den1 den2
  double mul, dec;
};

__host__ __device__ __forceinline__ void ramp_step(BiquadCoeffState& c) {
  if (c.mul > 0.0) {
#ifdef __CUDA_ARCH__
    c.mul = __dsub_rn(c.mul, __dmul_rn(c.dec, 1.0));
    c.mul = c.mul > 0.0 ? c.mul : 0.0;
#pragma unroll
    for (int k = 0; k < 5; k++) c.cur[k] = __dsub_rn(c.tgt[k], __dmul_rn(c.mul, c.dif[k]));
#else
    c.mul -= c.dec * 1.0;
    c.mul = c.mul > 0.0 ? c.mul : 0.0;
    for (int k = 0; k < 5; k++) {
      volatile double p = c.mul * c.dif[k];  // rounded product, then rounded difference (no contraction)
      c.cur[k] = c.tgt[k] - p;
    }
#endif
  }
}

// one thread per channel; frames in chunks of 8 so that the loads of a chunk are in flight together
__global__ void __launch_bounds__(128) k_biquad(const float* __restrict__ src, float* __restrict__ dst, double* __restrict__ w,
                                               BiquadCoeffState c, uint32_t nchannels, uint32_t nsrc, uint32_t ndst,
                                               uint32_t nframes) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nchannels) return;
  double w0 = w[2 * (size_t)j], w1 = w[2 * (size_t)j + 1];
  for (uint32_t i0 = 0; i0 < nframes; i0 += 8) {
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; u++) x[u] = (i0 + u < nframes) ? src[(size_t)(i0 + u) * nsrc + j] : 0.f;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (i0 + u < nframes) {
        const double xd = (double)x[u];
        const float y = __double2float_rn(__dadd_rn(__dmul_rn(xd, c.cur[0]), w0));
        const double yd = (double)y;
        w0 = __dadd_rn(__dsub_rn(__dmul_rn(xd, c.cur[1]), __dmul_rn(yd, c.cur[3])), w1);
        w1 = __dsub_rn(__dmul_rn(xd, c.cur[2]), __dmul_rn(yd, c.cur[4]));
        dst[(size_t)(i0 + u) * ndst + j] = y;
        ramp_step(c);
      }
    }
  }
  w[2 * (size_t)j] = w0;
  w[2 * (size_t)j + 1] = w1;
}

// ---- BiQuadFilterBank (src/BiQuad.h:247-353, src/BiQuad.cpp:498-690) ---------------------------------------------------------
// The reference runs the bank filter by filter: every filter makes a full pass over the block (BiQuad::Process with that
// filter's coefficient ramp), the first from src to dst, the others in place on dst (src/BiQuad.cpp:639-662).  Channel j of
// filter i only ever sees channel j of filter i - 1 at the same frame, so the passes fuse: one thread per channel takes a
// frame through ALL filters of the bank before it touches the next frame -- one read of src and one write of dst instead of
// one read and one write per filter, with the same IEEE operations in the same order per (filter, channel): bit-exact.
// Up to kFbankMax filters per launch (their coefficient states travel as kernel arguments); longer banks run in passes of
// kFbankMax, the later ones in place like the reference's.  While a filter ramps, its coefficients differ per frame: the CTA
// computes the trajectories of a chunk of kFbankChunk frames once (thread s = filter s, the reference's own recurrence) into
// shared memory and every channel reads them from there.
static constexpr int kFbankMax = 16, kFbankChunk = 8, kFbankThreads = 64;

struct FbankArgs {
  BiquadCoeffState c[kFbankMax];
};

template <int NST>
__global__ void __launch_bounds__(kFbankThreads) k_fbank(const float* __restrict__ src, float* __restrict__ dst, double* __restrict__ w,
                                                        const __grid_constant__ FbankArgs a, uint32_t nst, uint32_t wstride,
                                                        uint32_t nchannels, uint32_t nsrc, uint32_t ndst, uint32_t nframes,
                                                        int ramping) {
  __shared__ double coef[kFbankChunk][NST][5];
  __shared__ BiquadCoeffState ramp[NST];
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < nchannels;
  if (threadIdx.x < nst) {
    ramp[threadIdx.x] = a.c[threadIdx.x];
#pragma unroll
    for (int k = 0; k < 5; k++) coef[0][threadIdx.x][k] = a.c[threadIdx.x].cur[k];
  }
  double w0[NST], w1[NST];
#pragma unroll
  for (int s = 0; s < NST; s++) {
    const bool on = live && (uint32_t)s < nst;
    w0[s] = on ? w[((size_t)s * wstride + j) * 2] : 0.0;
    w1[s] = on ? w[((size_t)s * wstride + j) * 2 + 1] : 0.0;
  }
  __syncthreads();
  for (uint32_t i0 = 0; i0 < nframes; i0 += kFbankChunk) {
    if (ramping) {
      if (i0) __syncthreads();  // everybody is through with the previous chunk's coefficients
      if (threadIdx.x < nst) {
        BiquadCoeffState c = ramp[threadIdx.x];
        for (int u = 0; u < kFbankChunk; u++) {
#pragma unroll
          for (int k = 0; k < 5; k++) coef[u][threadIdx.x][k] = c.cur[k];
          ramp_step(c);
        }
        ramp[threadIdx.x] = c;
      }
      __syncthreads();
    }
    float x[kFbankChunk];
#pragma unroll
    for (int u = 0; u < kFbankChunk; u++) x[u] = (live && i0 + u < nframes) ? src[(size_t)(i0 + u) * nsrc + j] : 0.f;
#pragma unroll
    for (int u = 0; u < kFbankChunk; u++) {
      if (live && i0 + u < nframes) {
        float v = x[u];
#pragma unroll
        for (int s = 0; s < NST; s++) {
          if ((uint32_t)s < nst) {
            const double* c = coef[ramping ? u : 0][s];
            const double xd = (double)v;
            const float y = __double2float_rn(__dadd_rn(__dmul_rn(xd, c[0]), w0[s]));
            const double yd = (double)y;
            w0[s] = __dadd_rn(__dsub_rn(__dmul_rn(xd, c[1]), __dmul_rn(yd, c[3])), w1[s]);
            w1[s] = __dsub_rn(__dmul_rn(xd, c[2]), __dmul_rn(yd, c[4]));
            v = y;
          }
        }
        dst[(size_t)(i0 + u) * ndst + j] = v;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < NST; s++) {
    if (live && (uint32_t)s < nst) {
      w[((size_t)s * wstride + j) * 2] = w0[s];
      w[((size_t)s * wstride + j) * 2 + 1] = w1[s];
    }
  }
}

static void flat_coeffs(BiquadCoeffState& c) {
  memset(&c, 0, sizeof(c));
  c.cur[0] = c.tgt[0] = 1.0;  // BiQuadCoeffs(): flat, mul 0, dec 1 (src/BiQuad.cpp:11-24)
  c.dec = 1.0;
}

// target coefficients of a filter description, normalised by a0 (src/BiQuad.cpp:181-330; types src/BiQuad.h:31-42)
static void design(int type, double freq, double fs, double gain, double bandwidth, double* t) {
  const double A = pow(10.0, gain / 40.0);
  const double omega = 2.0 * M_PI * freq / fs;
  const double sn = sin(omega), cs = cos(omega);
  const double alpha = sn * sinh(M_LN2 / 2.0 * bandwidth * omega / sn);
  const double beta = sqrt(A + A);
  double b0 = 1.0, b1 = 0.0, b2 = 0.0, a0 = 1.0, a1 = 0.0, a2 = 0.0;
  switch (type) {
    case BBX_BIQUAD_LPF6: b0 = sn, a0 = 1 + sn, a1 = -1; break;
    case BBX_BIQUAD_LPF12: b0 = sn * sn, a0 = (1 + sn) * (1 + sn), a1 = -2 * (1 + sn), a2 = 1; break;
    case BBX_BIQUAD_HPF6: b0 = 1, b1 = -1, a1 = -(1 - sn); break;
    case BBX_BIQUAD_HPF12: b0 = 1, b1 = -2, b2 = 1, a1 = -2 * (1 - sn), a2 = (1 - sn) * (1 - sn); break;
    case BBX_BIQUAD_BPF: b0 = alpha, b2 = -alpha, a0 = 1 + alpha, a1 = -2 * cs, a2 = 1 - alpha; break;
    case BBX_BIQUAD_NOTCH: b0 = 1, b1 = -2 * cs, b2 = 1, a0 = 1 + alpha, a1 = -2 * cs, a2 = 1 - alpha; break;
    case BBX_BIQUAD_PEQ:
      b0 = 1 + (alpha * A), b1 = -2 * cs, b2 = 1 - (alpha * A);
      a0 = 1 + (alpha / A), a1 = -2 * cs, a2 = 1 - (alpha / A);
      break;
    case BBX_BIQUAD_LSH:
      b0 = A * ((A + 1) - (A - 1) * cs + beta * sn);
      b1 = 2 * A * ((A - 1) - (A + 1) * cs);
      b2 = A * ((A + 1) - (A - 1) * cs - beta * sn);
      a0 = (A + 1) + (A - 1) * cs + beta * sn;
      a1 = -2 * ((A - 1) + (A + 1) * cs);
      a2 = (A + 1) + (A - 1) * cs - beta * sn;
      break;
    case BBX_BIQUAD_HSH:
      b0 = A * ((A + 1) + (A - 1) * cs + beta * sn);
      b1 = -2 * A * ((A - 1) + (A + 1) * cs);
      b2 = A * ((A + 1) + (A - 1) * cs - beta * sn);
      a0 = (A + 1) - (A - 1) * cs + beta * sn;
      a1 = 2 * ((A - 1) - (A + 1) * cs);
      a2 = (A + 1) - (A - 1) * cs - beta * sn;
      break;
    default: break;  // FLAT
  }
  const double normalise = 1.0 / a0;
  t[0] = b0 * normalise;
  t[1] = b1 * normalise;
  t[2] = b2 * normalise;
  t[3] = a1 * normalise;
  t[4] = a2 * normalise;
}

static void retarget(BiquadCoeffState& c, double steps) {
  for (int i = 0; i < 5; i++) c.dif[i] = c.tgt[i] - c.cur[i];
  if (steps > 0.0) {
    c.mul = 1.0;
    c.dec = 1.0 / steps;
  } else {
    c.mul = c.dec = 0.0;
    memcpy(c.cur, c.tgt, sizeof(c.cur));
  }
}

}  // namespace bbx

using namespace bbx;

struct bbx_biquad {
  uint32_t nch = 0;
  int device = 0;
  BiquadCoeffState c;
  double* w = nullptr;  // device [nch][2]
};

// BiQuadFilterBank: per filter one coefficient object (host) and one state pair per channel (device, [filter][channel][2])
struct bbx_fbank {
  uint32_t nch = 0;
  int device = 0;
  std::vector<BiquadCoeffState> c;
  double* w = nullptr;
  uint64_t launches = 0;
};

// SetFilters / SetChannels (src/BiQuad.cpp:528-600): filters are dropped from / appended at the end, existing (filter,
// channel) pairs keep their audio state, new ones start from zero with flat coefficients
static int fbank_resize(bbx_fbank* f, uint32_t nch, uint32_t nfilters) {
  const uint32_t och = f->nch, ofl = (uint32_t)f->c.size();
  if (nch == och && nfilters == ofl) return BBX_OK;
  double* nw = nullptr;
  const size_t bytes = sizeof(double) * 2 * (size_t)nch * nfilters;
  if (bytes) {
    BBX_CUDA_TRY(cudaMalloc((void**)&nw, bytes));
    cudaError_t err = cudaMemset(nw, 0, bytes);
    const uint32_t kch = std::min(nch, och), kfl = std::min(nfilters, ofl);
    if (err == cudaSuccess && f->w && kch && kfl) {
      BBX_CUDA_TRY(cudaStreamSynchronize(cudaStreamPerThread));
      err = cudaMemcpy2D(nw, sizeof(double) * 2 * nch, f->w, sizeof(double) * 2 * och, sizeof(double) * 2 * kch, kfl,
                         cudaMemcpyDeviceToDevice);
    }
    if (err != cudaSuccess) {
      cudaFree(nw);
      BBX_CUDA_TRY(err);
    }
  }
  cudaFree(f->w);
  f->w = nw;
  f->nch = nch;
  BiquadCoeffState flat;
  flat_coeffs(flat);
  f->c.resize(nfilters, flat);
  return BBX_OK;
}

extern "C" {

int bbx_biquad_calc_coeffs(int type, double freq, double fs, double gain, double bandwidth, double* out5) {
  BBX_REQUIRE(out5 != nullptr, "bbx_biquad_calc_coeffs: null output");
  design(type, freq, fs, gain, bandwidth, out5);
  return BBX_OK;
}

int bbx_biquad_create(uint32_t nchannels, bbx_biquad** out) {
  BBX_REQUIRE(out != nullptr, "bbx_biquad_create: null argument");
  int rc = require_device();
  if (rc) return rc;
  bbx_biquad* b = new bbx_biquad();
  CreateGuard<bbx_biquad> guard(b, bbx_biquad_destroy);
  b->nch = nchannels;
  BBX_CUDA_TRY(cudaGetDevice(&b->device));
  memset(&b->c, 0, sizeof(b->c));
  b->c.cur[0] = b->c.tgt[0] = 1.0;  // BiQuadCoeffs(): flat, mul 0, dec 1 (src/BiQuad.cpp:11-24)
  b->c.dec = 1.0;
  BBX_CUDA_TRY(cudaMalloc((void**)&b->w, sizeof(double) * 2 * (size_t)(nchannels ? nchannels : 1)));
  BBX_CUDA_TRY(cudaMemset(b->w, 0, sizeof(double) * 2 * (size_t)(nchannels ? nchannels : 1)));
  *out = guard.release();
  return BBX_OK;
}

int bbx_biquad_destroy(bbx_biquad* b) {
  if (!b) return BBX_OK;
  DeviceGuard dg(b->device);
  cudaFree(b->w);
  delete b;
  return BBX_OK;
}

int bbx_biquad_set_coeffs(bbx_biquad* b, const double* c5, double interp_samples) {
  BBX_REQUIRE(b && c5, "bbx_biquad_set_coeffs: null argument");
  memcpy(b->c.tgt, c5, sizeof(b->c.tgt));
  retarget(b->c, interp_samples);
  return BBX_OK;
}

int bbx_biquad_calc(bbx_biquad* b, int type, double freq, double fs, double gain, double bandwidth, double interp_time) {
  BBX_REQUIRE(b != nullptr, "bbx_biquad_calc: null argument");
  design(type, freq, fs, gain, bandwidth, b->c.tgt);
  retarget(b->c, interp_time > 0.0 ? interp_time * fs : 0.0);
  return BBX_OK;
}

int bbx_biquad_process_dev(bbx_biquad* b, const float* src, float* dst, uint32_t nchannels, uint32_t nsrcchannels,
                           uint32_t ndstchannels, uint32_t nframes, void* stream) {
  BBX_REQUIRE(b != nullptr, "bbx_biquad_process: null argument");
  nchannels = std::min(std::min(nchannels, nsrcchannels), std::min(ndstchannels, b->nch));
  if (!nchannels || !nframes) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_biquad_process: null buffer");
  DeviceGuard dg(b->device);
  k_biquad<<<ceil_div(nchannels, 128u), 128, 0, (cudaStream_t)stream>>>(src, dst, b->w, b->c, nchannels, nsrcchannels,
                                                                        ndstchannels, nframes);
  BBX_CUDA_TRY(cudaGetLastError());
  // the caller's coefficient object advances one ramp step per frame (host copy of the same recurrence)
  for (uint32_t i = 0; i < nframes && b->c.mul > 0.0; i++) ramp_step(b->c);
  return BBX_OK;
}

int bbx_biquad_process(bbx_biquad* b, const float* src, float* dst, uint32_t nchannels, uint32_t nsrcchannels,
                       uint32_t ndstchannels, uint32_t nframes) {
  BBX_REQUIRE(b != nullptr, "bbx_biquad_process: null argument");
  if (!nframes || !nsrcchannels || !ndstchannels) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_biquad_process: null buffer");
  const size_t sb = sizeof(float) * (size_t)nframes * nsrcchannels, db = sizeof(float) * (size_t)nframes * ndstchannels;
  DeviceGuard dg(b->device);
  DeviceScratch& s0 = scratch(0);
  DeviceScratch& s1 = scratch(1);
  int rc;
  if ((rc = s0.ensure(sb)) || (rc = s1.ensure(db))) return rc;
  cudaStream_t st = cudaStreamPerThread;
  BBX_CUDA_TRY(cudaMemcpyAsync(s0.ptr, src, sb, cudaMemcpyHostToDevice, st));
  // channels beyond nchannels keep the caller's dst bytes
  BBX_CUDA_TRY(cudaMemcpyAsync(s1.ptr, dst, db, cudaMemcpyHostToDevice, st));
  if ((rc = bbx_biquad_process_dev(b, (const float*)s0.ptr, (float*)s1.ptr, nchannels, nsrcchannels, ndstchannels, nframes, st)))
    return rc;
  BBX_CUDA_TRY(cudaMemcpyAsync(dst, s1.ptr, db, cudaMemcpyDeviceToHost, st));
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  return BBX_OK;
}

int bbx_biquad_get_state(const bbx_biquad* b, double* w, double* cur5, double* mul_dec) {
  BBX_REQUIRE(b != nullptr, "bbx_biquad_get_state: null argument");
  DeviceGuard dg(b->device);
  if (w && b->nch) {
    BBX_CUDA_TRY(cudaStreamSynchronize(cudaStreamPerThread));
    BBX_CUDA_TRY(cudaMemcpy(w, b->w, sizeof(double) * 2 * (size_t)b->nch, cudaMemcpyDeviceToHost));
  }
  if (cur5) memcpy(cur5, b->c.cur, sizeof(b->c.cur));
  if (mul_dec) {
    mul_dec[0] = b->c.mul;
    mul_dec[1] = b->c.dec;
  }
  return BBX_OK;
}

int bbx_biquad_reset(bbx_biquad* b) {
  BBX_REQUIRE(b != nullptr, "bbx_biquad_reset: null argument");
  DeviceGuard dg(b->device);
  BBX_CUDA_TRY(cudaMemset(b->w, 0, sizeof(double) * 2 * (size_t)(b->nch ? b->nch : 1)));
  return BBX_OK;
}

/* ---- BiQuadFilterBank ---- */
int bbx_fbank_create(uint32_t nchannels, uint32_t nfilters, bbx_fbank** out) {
  BBX_REQUIRE(out != nullptr, "bbx_fbank_create: null argument");
  int rc = require_device();
  if (rc) return rc;
  bbx_fbank* f = new bbx_fbank();
  CreateGuard<bbx_fbank> guard(f, bbx_fbank_destroy);
  BBX_CUDA_TRY(cudaGetDevice(&f->device));
  if ((rc = fbank_resize(f, nchannels, nfilters))) return rc;
  *out = guard.release();
  return BBX_OK;
}

int bbx_fbank_destroy(bbx_fbank* f) {
  if (!f) return BBX_OK;
  DeviceGuard dg(f->device);
  cudaFree(f->w);
  delete f;
  return BBX_OK;
}

int bbx_fbank_set_filters(bbx_fbank* f, uint32_t n) {
  BBX_REQUIRE(f != nullptr, "bbx_fbank_set_filters: null argument");
  DeviceGuard dg(f->device);
  return fbank_resize(f, f->nch, n);
}

int bbx_fbank_add_filter(bbx_fbank* f, const double* c5) {
  BBX_REQUIRE(f && c5, "bbx_fbank_add_filter: null argument");
  DeviceGuard dg(f->device);
  int rc = fbank_resize(f, f->nch, (uint32_t)f->c.size() + 1);
  if (rc) return rc;
  memcpy(f->c.back().tgt, c5, sizeof(f->c.back().tgt));
  retarget(f->c.back(), 0.0);
  return BBX_OK;
}

int bbx_fbank_set_channels(bbx_fbank* f, uint32_t n) {
  BBX_REQUIRE(f != nullptr, "bbx_fbank_set_channels: null argument");
  DeviceGuard dg(f->device);
  return fbank_resize(f, n, (uint32_t)f->c.size());
}

int bbx_fbank_get_size(const bbx_fbank* f, uint32_t* nchannels, uint32_t* nfilters) {
  BBX_REQUIRE(f != nullptr, "bbx_fbank_get_size: null argument");
  if (nchannels) *nchannels = f->nch;
  if (nfilters) *nfilters = (uint32_t)f->c.size();
  return BBX_OK;
}

int bbx_fbank_set_coeffs(bbx_fbank* f, uint32_t filter, const double* c5, double interp_samples) {
  BBX_REQUIRE(f && c5, "bbx_fbank_set_coeffs: null argument");
  BBX_REQUIRE(filter < f->c.size(), "bbx_fbank_set_coeffs: filter %u of %zu", filter, f->c.size());
  memcpy(f->c[filter].tgt, c5, sizeof(f->c[filter].tgt));
  retarget(f->c[filter], interp_samples);
  return BBX_OK;
}

int bbx_fbank_calc(bbx_fbank* f, uint32_t filter, int type, double freq, double fs, double gain, double bandwidth,
                   double interp_time) {
  BBX_REQUIRE(f != nullptr, "bbx_fbank_calc: null argument");
  BBX_REQUIRE(filter < f->c.size(), "bbx_fbank_calc: filter %u of %zu", filter, f->c.size());
  design(type, freq, fs, gain, bandwidth, f->c[filter].tgt);
  retarget(f->c[filter], interp_time > 0.0 ? interp_time * fs : 0.0);
  return BBX_OK;
}

int bbx_fbank_process_dev(bbx_fbank* f, const float* src, float* dst, uint32_t nchannels, uint32_t nsrcchannels,
                          uint32_t ndstchannels, uint32_t nframes, void* stream) {
  BBX_REQUIRE(f != nullptr, "bbx_fbank_process: null argument");
  nchannels = std::min(std::min(nchannels, nsrcchannels), std::min(ndstchannels, f->nch));
  if (!nchannels || !nframes || f->c.empty()) return BBX_OK;  // a bank without filters leaves dst alone (src/BiQuad.cpp:645)
  BBX_REQUIRE(src && dst, "bbx_fbank_process: null buffer");
  DeviceGuard dg(f->device);
  const uint32_t total = (uint32_t)f->c.size();
  for (uint32_t s0 = 0; s0 < total; s0 += kFbankMax) {
    const uint32_t nst = std::min<uint32_t>(kFbankMax, total - s0);
    FbankArgs a;
    int ramping = 0;
    for (uint32_t s = 0; s < nst; s++) {
      a.c[s] = f->c[s0 + s];
      ramping |= a.c[s].mul > 0.0;
    }
    for (uint32_t s = nst; s < kFbankMax; s++) flat_coeffs(a.c[s]);
    double* w = f->w + (size_t)s0 * f->nch * 2;
    const dim3 grid(ceil_div(nchannels, (uint32_t)kFbankThreads));
    const cudaStream_t st = (cudaStream_t)stream;
    // the later passes of a long bank run in place on dst, like the reference's later filters
    const float* in = s0 ? dst : src;
    const uint32_t nin = s0 ? ndstchannels : nsrcchannels;
    if (nst <= 2)
      k_fbank<2><<<grid, kFbankThreads, 0, st>>>(in, dst, w, a, nst, f->nch, nchannels, nin, ndstchannels, nframes, ramping);
    else if (nst <= 4)
      k_fbank<4><<<grid, kFbankThreads, 0, st>>>(in, dst, w, a, nst, f->nch, nchannels, nin, ndstchannels, nframes, ramping);
    else if (nst <= 8)
      k_fbank<8><<<grid, kFbankThreads, 0, st>>>(in, dst, w, a, nst, f->nch, nchannels, nin, ndstchannels, nframes, ramping);
    else
      k_fbank<16><<<grid, kFbankThreads, 0, st>>>(in, dst, w, a, nst, f->nch, nchannels, nin, ndstchannels, nframes, ramping);
    BBX_CUDA_TRY(cudaGetLastError());
    f->launches++;
  }
  // the coefficient objects advance one ramp step per frame (host copy of the same recurrence)
  for (BiquadCoeffState& c : f->c)
    for (uint32_t i = 0; i < nframes && c.mul > 0.0; i++) ramp_step(c);
  return BBX_OK;
}

int bbx_fbank_process(bbx_fbank* f, const float* src, float* dst, uint32_t nchannels, uint32_t nsrcchannels,
                      uint32_t ndstchannels, uint32_t nframes) {
  BBX_REQUIRE(f != nullptr, "bbx_fbank_process: null argument");
  if (!nframes || !nsrcchannels || !ndstchannels) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_fbank_process: null buffer");
  const size_t sb = sizeof(float) * (size_t)nframes * nsrcchannels, db = sizeof(float) * (size_t)nframes * ndstchannels;
  DeviceGuard dg(f->device);
  DeviceScratch& s0 = scratch(0);
  DeviceScratch& s1 = scratch(1);
  int rc;
  if ((rc = s0.ensure(sb)) || (rc = s1.ensure(db))) return rc;
  cudaStream_t st = cudaStreamPerThread;
  BBX_CUDA_TRY(cudaMemcpyAsync(s0.ptr, src, sb, cudaMemcpyHostToDevice, st));
  BBX_CUDA_TRY(cudaMemcpyAsync(s1.ptr, dst, db, cudaMemcpyHostToDevice, st));  // channels beyond nchannels keep the caller's bytes
  if ((rc = bbx_fbank_process_dev(f, (const float*)s0.ptr, (float*)s1.ptr, nchannels, nsrcchannels, ndstchannels, nframes, st)))
    return rc;
  BBX_CUDA_TRY(cudaMemcpyAsync(dst, s1.ptr, db, cudaMemcpyDeviceToHost, st));
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  return BBX_OK;
}

int bbx_fbank_get_state(const bbx_fbank* f, uint32_t filter, double* w, double* cur5, double* mul_dec) {
  BBX_REQUIRE(f != nullptr, "bbx_fbank_get_state: null argument");
  BBX_REQUIRE(filter < f->c.size(), "bbx_fbank_get_state: filter %u of %zu", filter, f->c.size());
  DeviceGuard dg(f->device);
  if (w && f->nch) {
    BBX_CUDA_TRY(cudaStreamSynchronize(cudaStreamPerThread));
    BBX_CUDA_TRY(cudaMemcpy(w, f->w + (size_t)filter * f->nch * 2, sizeof(double) * 2 * (size_t)f->nch, cudaMemcpyDeviceToHost));
  }
  if (cur5) memcpy(cur5, f->c[filter].cur, sizeof(f->c[filter].cur));
  if (mul_dec) {
    mul_dec[0] = f->c[filter].mul;
    mul_dec[1] = f->c[filter].dec;
  }
  return BBX_OK;
}

int bbx_fbank_reset(bbx_fbank* f) {
  BBX_REQUIRE(f != nullptr, "bbx_fbank_reset: null argument");
  DeviceGuard dg(f->device);
  if (f->w) BBX_CUDA_TRY(cudaMemset(f->w, 0, sizeof(double) * 2 * (size_t)f->nch * f->c.size()));
  return BBX_OK;
}

int bbx_fbank_launches(const bbx_fbank* f, uint64_t* launches) {
  BBX_REQUIRE(f && launches, "bbx_fbank_launches: null argument");
  *launches = f->launches;
  return BBX_OK;
}

}  // extern "C"
