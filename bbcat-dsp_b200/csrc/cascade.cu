// cascade.cu -- BiQuadCascade on the GPU as a bank of independent cascades, one per channel (the EQ stage
// that follows the convolver in a renderer chain).  Replaces, per channel,
//   BiQuadCascade::SetCoefficients(interleaved vector) / Reset / ProcessCascade(input, dest, blocksize)   src/BiQuad.h:531-558, :517-524, :718-739
// with both forms of Tick (src/BiQuad.h:682-716): the plain cascade (filter i reads y[i-1] of the same sample) and the
// "vectorised" pipeline the SSE3 build runs (every filter reads the x register its predecessor wrote on the previous
// sample; x[1..] = y[0..] after each sample; numfilters must be a multiple of four, else the reference switches it off).
// The recurrence is serial in time and at most 12 filters deep, independent across channels -> one thread per channel
// walks the frames with the 4 x 12 registers and 4 x 12 coefficients in its register file.  float arithmetic with every
// product, difference and sum rounded separately (__fmul_rn / __fsub_rn / __fadd_rn, the order of src/BiQuad.h:667-672 and
// of the intrinsics at :600-625): bit-exact against the reference build.  The output gain g is stored and, as in the
// reference, never applied.
#include <vector>

#include "common.cuh"

namespace bbx {

constexpr int kCascMax = 12;  // BiQuadCascade::maxnumfilters (src/BiQuad.h:772)

struct CascadeRegs {  // one channel, the reference's member order
  float b1[kCascMax], b2[kCascMax], a1[kCascMax], a2[kCascMax];
  float x[kCascMax], y[kCascMax], w0[kCascMax], w1[kCascMax];
  float lastoutput, g;
};

template <bool VEC>
__global__ void __launch_bounds__(64) k_cascade(CascadeRegs* __restrict__ regs, uint32_t nch, uint32_t nf, const float* src,
                                                long long src_cs, long long src_fs, float* dst, long long dst_cs,
                                                long long dst_fs, uint32_t nframes) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nch) return;
  CascadeRegs& r = regs[j];
  float b1[kCascMax], b2[kCascMax], a1[kCascMax], a2[kCascMax], x[kCascMax], y[kCascMax], w0[kCascMax], w1[kCascMax];
#pragma unroll
  for (int i = 0; i < kCascMax; i++) {
    b1[i] = r.b1[i]; b2[i] = r.b2[i]; a1[i] = r.a1[i]; a2[i] = r.a2[i];
    x[i] = r.x[i]; y[i] = r.y[i]; w0[i] = r.w0[i]; w1[i] = r.w1[i];
  }
  const float* in = src + (long long)j * src_cs;
  float* out = dst + (long long)j * dst_cs;
  float last = r.lastoutput;
  for (uint32_t n = 0; n < nframes; n++) {
    const float v = in[(long long)n * src_fs];
    if (VEC) {
      x[0] = v;
#pragma unroll
      for (int i = 0; i < kCascMax; i++)
        if ((uint32_t)i < nf) {
          const float yv = __fadd_rn(x[i], w0[i]);
          y[i] = yv;
          w0[i] = __fadd_rn(__fsub_rn(__fmul_rn(x[i], b1[i]), __fmul_rn(yv, a1[i])), w1[i]);
          w1[i] = __fsub_rn(__fmul_rn(x[i], b2[i]), __fmul_rn(yv, a2[i]));
        }
      // copy filter outputs to filter inputs: x[1 .. nf-1] = y[0 .. nf-2]
#pragma unroll
      for (int i = kCascMax - 1; i >= 1; i--)
        if ((uint32_t)i < nf) x[i] = y[i - 1];
    } else {
      float p = v;  // input of filter i: the sample, then y[i-1]
#pragma unroll
      for (int i = 0; i < kCascMax; i++)
        if ((uint32_t)i < nf) {
          const float yv = __fadd_rn(p, w0[i]);
          y[i] = yv;
          w0[i] = __fadd_rn(__fsub_rn(__fmul_rn(p, b1[i]), __fmul_rn(yv, a1[i])), w1[i]);
          w1[i] = __fsub_rn(__fmul_rn(p, b2[i]), __fmul_rn(yv, a2[i]));
          p = yv;
        }
    }
    // y[nf - 1]
    float o = y[0];
#pragma unroll
    for (int i = 1; i < kCascMax; i++)
      if ((uint32_t)i == nf - 1) o = y[i];
    last = o;
    out[(long long)n * dst_fs] = o;
  }
#pragma unroll
  for (int i = 0; i < kCascMax; i++) {
    r.x[i] = x[i]; r.y[i] = y[i]; r.w0[i] = w0[i]; r.w1[i] = w1[i];
  }
  r.lastoutput = last;
}

}  // namespace bbx

using namespace bbx;

struct bbx_cascade {
  uint32_t nch = 0, nf = 0;
  bool vectorise = false;
  int device = 0;
  CascadeRegs* d_regs = nullptr;
};

extern "C" {

int bbx_cascade_create(uint32_t nchannels, uint32_t numfilters, int vectorise, int unroll, bbx_cascade** out) {
  (void)unroll;  // BiQuadCascade::ProcessCascade: the unrolled loop is the same arithmetic (src/BiQuad.h:720-738)
  BBX_REQUIRE(out && nchannels >= 1, "bbx_cascade_create: bad argument");
  // the reference logs an error and leaves a cascade of 0 filters whose Tick is undefined (src/BiQuad.h:400-404, :696-699)
  BBX_REQUIRE(numfilters >= 1 && numfilters <= (uint32_t)kCascMax, "bbx_cascade_create: numfilters %u outside 1..%d", numfilters,
              kCascMax);
  int rc = require_device();
  if (rc) return rc;
  bbx_cascade* c = new bbx_cascade();
  CreateGuard<bbx_cascade> guard(c, bbx_cascade_destroy);
  c->nch = nchannels;
  c->nf = numfilters;
  BBX_CUDA_TRY(cudaGetDevice(&c->device));
  c->vectorise = vectorise && (numfilters % 4) == 0;  // src/BiQuad.h:405-409
  BBX_CUDA_TRY(cudaMalloc((void**)&c->d_regs, sizeof(CascadeRegs) * nchannels));
  // pass-through default: zero coefficients, g = 1 (src/BiQuad.h:410-415); registers start from Reset()
  std::vector<CascadeRegs> h(nchannels);
  memset(h.data(), 0, sizeof(CascadeRegs) * nchannels);
  for (auto& r : h) r.g = 1.0f;
  BBX_CUDA_TRY(cudaMemcpy(c->d_regs, h.data(), sizeof(CascadeRegs) * nchannels, cudaMemcpyHostToDevice));
  *out = guard.release();
  return BBX_OK;
}

int bbx_cascade_destroy(bbx_cascade* c) {
  if (!c) return BBX_OK;
  DeviceGuard dg(c->device);
  cudaFree(c->d_regs);
  delete c;
  return BBX_OK;
}

int bbx_cascade_set_coefficients(bbx_cascade* c, uint32_t channel, const float* coeffs, uint32_t n) {
  BBX_REQUIRE(c && coeffs, "bbx_cascade_set_coefficients: null argument");
  // "coefficients vector must be 4*numfilters + 1 long" (src/BiQuad.h:533-537)
  BBX_REQUIRE(n == 4 * c->nf + 1, "bbx_cascade_set_coefficients: %u coefficients, expected 4 * %u + 1", n, c->nf);
  BBX_REQUIRE(channel == 0xFFFFFFFFu || channel < c->nch, "bbx_cascade_set_coefficients: channel %u outside the bank", channel);
  DeviceGuard dg(c->device);
  CascadeRegs r;
  memset(&r, 0, sizeof(r));  // SetCoefficients ends with Reset()
  const float* p = coeffs;
  r.g = *p++;
  for (uint32_t i = 0; i < c->nf; i++) {
    r.b1[i] = *p++;
    r.b2[i] = *p++;
    r.a1[i] = *p++;
    r.a2[i] = *p++;
  }
  BBX_CUDA_TRY(cudaStreamSynchronize(cudaStreamPerThread));
  if (channel == 0xFFFFFFFFu) {
    std::vector<CascadeRegs> h(c->nch, r);
    BBX_CUDA_TRY(cudaMemcpy(c->d_regs, h.data(), sizeof(CascadeRegs) * c->nch, cudaMemcpyHostToDevice));
  } else {
    // filters beyond numfilters keep their (unused) coefficients, as in the reference
    CascadeRegs cur;
    BBX_CUDA_TRY(cudaMemcpy(&cur, c->d_regs + channel, sizeof(cur), cudaMemcpyDeviceToHost));
    for (uint32_t i = c->nf; i < (uint32_t)kCascMax; i++) {
      r.b1[i] = cur.b1[i];
      r.b2[i] = cur.b2[i];
      r.a1[i] = cur.a1[i];
      r.a2[i] = cur.a2[i];
    }
    BBX_CUDA_TRY(cudaMemcpy(c->d_regs + channel, &r, sizeof(r), cudaMemcpyHostToDevice));
  }
  return BBX_OK;
}

int bbx_cascade_reset(bbx_cascade* c) {
  BBX_REQUIRE(c != nullptr, "bbx_cascade_reset: null argument");
  DeviceGuard dg(c->device);
  BBX_CUDA_TRY(cudaStreamSynchronize(cudaStreamPerThread));
  // x, y, w0, w1, lastoutput are contiguous in CascadeRegs
  const size_t off = offsetof(CascadeRegs, x), len = offsetof(CascadeRegs, g) - off;
  BBX_CUDA_TRY(cudaMemset2D((uint8_t*)c->d_regs + off, sizeof(CascadeRegs), 0, len, c->nch));
  return BBX_OK;
}

int bbx_cascade_process_dev(bbx_cascade* c, const float* src, long long src_channel_stride, long long src_frame_stride,
                            float* dst, long long dst_channel_stride, long long dst_frame_stride, uint32_t nframes,
                            void* stream) {
  BBX_REQUIRE(c != nullptr, "bbx_cascade_process: null argument");
  if (!nframes) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_cascade_process: null buffer");
  DeviceGuard dg(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t grid = ceil_div(c->nch, 64u);
  if (c->vectorise)
    k_cascade<true><<<grid, 64, 0, st>>>(c->d_regs, c->nch, c->nf, src, src_channel_stride, src_frame_stride, dst,
                                         dst_channel_stride, dst_frame_stride, nframes);
  else
    k_cascade<false><<<grid, 64, 0, st>>>(c->d_regs, c->nch, c->nf, src, src_channel_stride, src_frame_stride, dst,
                                          dst_channel_stride, dst_frame_stride, nframes);
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

int bbx_cascade_process(bbx_cascade* c, const float* src, float* dst, uint32_t nframes, int interleaved) {
  BBX_REQUIRE(c != nullptr, "bbx_cascade_process: null argument");
  if (!nframes) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_cascade_process: null buffer");
  const size_t bytes = sizeof(float) * (size_t)nframes * c->nch;
  DeviceGuard dg(c->device);
  DeviceScratch& s0 = scratch(0);
  DeviceScratch& s1 = scratch(1);
  int rc;
  if ((rc = s0.ensure(bytes)) || (rc = s1.ensure(bytes))) return rc;
  cudaStream_t st = cudaStreamPerThread;
  BBX_CUDA_TRY(cudaMemcpyAsync(s0.ptr, src, bytes, cudaMemcpyHostToDevice, st));
  const long long cs = interleaved ? 1 : (long long)nframes, fs = interleaved ? (long long)c->nch : 1;
  if ((rc = bbx_cascade_process_dev(c, (const float*)s0.ptr, cs, fs, (float*)s1.ptr, cs, fs, nframes, st))) return rc;
  BBX_CUDA_TRY(cudaMemcpyAsync(dst, s1.ptr, bytes, cudaMemcpyDeviceToHost, st));
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  return BBX_OK;
}

uint32_t bbx_cascade_get_state(const bbx_cascade* c, uint32_t channel, float* x12, float* y12, float* w0_12, float* w1_12,
                               float* lastoutput) {
  if (!c || channel >= c->nch) return 0;
  CascadeRegs r;
  DeviceGuard dg(c->device);
  cudaStreamSynchronize(cudaStreamPerThread);
  if (cudaMemcpy(&r, c->d_regs + channel, sizeof(r), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  if (x12) memcpy(x12, r.x, sizeof(r.x));
  if (y12) memcpy(y12, r.y, sizeof(r.y));
  if (w0_12) memcpy(w0_12, r.w0, sizeof(r.w0));
  if (w1_12) memcpy(w1_12, r.w1, sizeof(r.w1));
  if (lastoutput) *lastoutput = r.lastoutput;
  return c->nf | ((uint32_t)c->vectorise << 8);
}

}  // extern "C"
