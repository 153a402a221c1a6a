// allpass.cu -- AllPassFilter<float> / AllPassFilterChain<float> on the GPU (decorrelation banks, the
// other "next" component after the convolver in a renderer chain).  Replaces
//   AllPassFilter::Process(src, dst, srcchannel, nsrcchannels, dstchannel, ndstchannels, nframes)   src/AllPassFilter.h:84-128
//   AllPassFilterChain::Process (section 0 reads src, the others run in place on dst)                src/AllPassFilter.h:238-255
// with the rings in HBM in the reference's own layout ([delay][nchannels] items, one position per section,
// src/RingBuffer.h:46-53).  A section is y[n] = c x[n] + w[n-d], w[n] = x[n] - c y[n]: serial in time with a dependency
// distance of d samples, independent across channels -> one thread per channel runs the whole chain, section by section.
// float arithmetic, products and sums rounded separately (__fmul_rn / __fadd_rn): bit-exact against the reference build.
#include <vector>

#include "common.cuh"

namespace bbx {

struct AllpassSection {
  float* ring;        // [delay][nchannels]
  uint32_t delay;
  uint32_t slot;      // ring position / nchannels at the start of the call
  float coeff;
  uint32_t pad;
};

__global__ void __launch_bounds__(128) k_allpass(const AllpassSection* __restrict__ sec, uint32_t nfilters, const float* src,
                                                 float* dst, uint32_t nch, uint32_t n, uint32_t srcchannel, uint32_t nsrc,
                                                 uint32_t dstchannel, uint32_t ndst, uint32_t nframes) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* in = src + srcchannel + j;
  uint32_t in_stride = nsrc;
  float* out = dst + dstchannel + j;
  for (uint32_t f = 0; f < nfilters; f++) {
    const AllpassSection s = sec[f];
    float* ring = s.ring + j;
    uint32_t slot = s.slot;
    const float c = s.coeff;
    for (uint32_t i = 0; i < nframes; i++) {
      const float x = in[(size_t)i * in_stride];
      const float wold = ring[(size_t)slot * nch];
      const float y = __fadd_rn(__fmul_rn(c, x), wold);
      ring[(size_t)slot * nch] = __fsub_rn(x, __fmul_rn(c, y));
      out[(size_t)i * ndst] = y;
      if (++slot >= s.delay) slot = 0;
    }
    in = out;  // the following sections run in place on dst
    in_stride = ndst;
  }
}

}  // namespace bbx

using namespace bbx;

struct bbx_allpass {
  uint32_t nch = 0, nf = 0;
  std::vector<uint32_t> delay, pos;  // pos in ring items, like RingBuffer::GetPosition()
  std::vector<float> coeff;
  std::vector<float*> ring;
  AllpassSection* d_sec = nullptr;
};

extern "C" {

int bbx_allpass_create(uint32_t nchannels, uint32_t nfilters, const uint32_t* delays, const float* coeffs, bbx_allpass** out) {
  BBX_REQUIRE(out && nchannels >= 1 && (nfilters == 0 || (delays && coeffs)), "bbx_allpass_create: bad argument");
  for (uint32_t f = 0; f < nfilters; f++)
    BBX_REQUIRE(delays[f] >= 1, "bbx_allpass_create: delay of section %u is 0 (the reference divides by the ring length)", f);
  int rc = require_device();
  if (rc) return rc;
  bbx_allpass* a = new bbx_allpass();
  CreateGuard<bbx_allpass> guard(a, bbx_allpass_destroy);
  a->nch = nchannels;
  a->nf = nfilters;
  a->delay.assign(delays, delays + nfilters);
  a->coeff.assign(coeffs, coeffs + nfilters);
  a->pos.assign(nfilters, 0);
  a->ring.assign(nfilters, nullptr);
  for (uint32_t f = 0; f < nfilters; f++) {
    const size_t bytes = sizeof(float) * (size_t)nchannels * delays[f];
    BBX_CUDA_TRY(cudaMalloc((void**)&a->ring[f], bytes));
    BBX_CUDA_TRY(cudaMemset(a->ring[f], 0, bytes));
  }
  BBX_CUDA_TRY(cudaMalloc((void**)&a->d_sec, sizeof(AllpassSection) * (nfilters ? nfilters : 1)));
  *out = guard.release();
  return BBX_OK;
}

int bbx_allpass_destroy(bbx_allpass* a) {
  if (!a) return BBX_OK;
  for (float* r : a->ring) cudaFree(r);
  cudaFree(a->d_sec);
  delete a;
  return BBX_OK;
}

int bbx_allpass_process_dev(bbx_allpass* a, const float* src, float* dst, uint32_t srcchannel, uint32_t nsrcchannels,
                            uint32_t dstchannel, uint32_t ndstchannels, uint32_t nframes, void* stream) {
  BBX_REQUIRE(a != nullptr, "bbx_allpass_process: null argument");
  if (!a->nf || !nframes) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_allpass_process: null buffer");
  // channels that fit both geometries (src/AllPassFilter.h:108-111); the single-channel branch does not clamp
  uint32_t n = a->nch;
  if (a->nch != 1) {
    n = std::min(n, nsrcchannels >= srcchannel ? nsrcchannels - srcchannel : 0u);
    n = std::min(n, ndstchannels >= dstchannel ? ndstchannels - dstchannel : 0u);
  } else {
    BBX_REQUIRE(srcchannel < nsrcchannels && dstchannel < ndstchannels, "bbx_allpass_process: channel outside the buffers");
  }
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<AllpassSection> sec(a->nf);
  for (uint32_t f = 0; f < a->nf; f++) {
    sec[f].ring = a->ring[f];
    sec[f].delay = a->delay[f];
    sec[f].slot = a->pos[f] / a->nch;
    sec[f].coeff = a->coeff[f];
    sec[f].pad = 0;
  }
  if (n) {
    BBX_CUDA_TRY(cudaMemcpyAsync(a->d_sec, sec.data(), sizeof(AllpassSection) * a->nf, cudaMemcpyHostToDevice, st));
    BBX_CUDA_TRY(cudaStreamSynchronize(st));  // sec is a stack vector
    k_allpass<<<ceil_div(n, 128u), 128, 0, st>>>(a->d_sec, a->nf, src, dst, a->nch, n, srcchannel, nsrcchannels, dstchannel,
                                                ndstchannels, nframes);
    BBX_CUDA_TRY(cudaGetLastError());
  }
  // every section's ring position advances nchannels items per frame, processed or skipped (Advance)
  for (uint32_t f = 0; f < a->nf; f++) {
    const uint64_t len = (uint64_t)a->nch * a->delay[f];
    a->pos[f] = (uint32_t)(((uint64_t)a->pos[f] + (uint64_t)nframes * a->nch) % len);
  }
  return BBX_OK;
}

int bbx_allpass_process(bbx_allpass* a, const float* src, float* dst, uint32_t srcchannel, uint32_t nsrcchannels,
                        uint32_t dstchannel, uint32_t ndstchannels, uint32_t nframes) {
  BBX_REQUIRE(a != nullptr, "bbx_allpass_process: null argument");
  if (!a->nf || !nframes || !nsrcchannels || !ndstchannels) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_allpass_process: null buffer");
  const size_t sb = sizeof(float) * (size_t)nframes * nsrcchannels, db = sizeof(float) * (size_t)nframes * ndstchannels;
  DeviceScratch& s0 = scratch(0);
  DeviceScratch& s1 = scratch(1);
  int rc;
  if ((rc = s0.ensure(sb)) || (rc = s1.ensure(db))) return rc;
  cudaStream_t st = cudaStreamPerThread;
  BBX_CUDA_TRY(cudaMemcpyAsync(s0.ptr, src, sb, cudaMemcpyHostToDevice, st));
  BBX_CUDA_TRY(cudaMemcpyAsync(s1.ptr, dst, db, cudaMemcpyHostToDevice, st));  // untouched channels keep the caller's values
  if ((rc = bbx_allpass_process_dev(a, (const float*)s0.ptr, (float*)s1.ptr, srcchannel, nsrcchannels, dstchannel, ndstchannels,
                                    nframes, st)))
    return rc;
  BBX_CUDA_TRY(cudaMemcpyAsync(dst, s1.ptr, db, cudaMemcpyDeviceToHost, st));
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  return BBX_OK;
}

uint32_t bbx_allpass_get_state(const bbx_allpass* a, uint32_t filter, float* ring, uint32_t maxitems) {
  if (!a || filter >= a->nf) return 0;
  uint32_t n = std::min(a->nch * a->delay[filter], maxitems);
  if (n && ring) {
    cudaStreamSynchronize(cudaStreamPerThread);
    cudaMemcpy(ring, a->ring[filter], sizeof(float) * n, cudaMemcpyDeviceToHost);
  }
  return a->pos[filter];
}

}  // extern "C"
