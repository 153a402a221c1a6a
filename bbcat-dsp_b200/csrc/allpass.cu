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

// up to kAllpassBatch sections travel by value in the launch parameters: the _dev entry point stays asynchronous (no
// staging copy, no stream synchronisation); longer chains are run as consecutive launches
constexpr uint32_t kAllpassBatch = 16;
struct AllpassBatch {
  AllpassSection sec[kAllpassBatch];
};

// n_first channels run the first section of the chain (it reads src), n_rest channels the following ones (in place on dst):
// the reference's chain switches to the dst geometry after its first filter and every filter recomputes its own channel
// count (src/AllPassFilter.h:108-111, 238-255), so a src narrower than dst leaves channels that only the later sections touch
__global__ void __launch_bounds__(128) k_allpass(const AllpassBatch b, uint32_t nfilters, uint32_t first_is_head, const float* src,
                                                 float* dst, uint32_t nch, uint32_t n_first, uint32_t n_rest, uint32_t srcchannel,
                                                 uint32_t nsrc, uint32_t dstchannel, uint32_t ndst, uint32_t nframes) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  float* out = dst + dstchannel + j;
  for (uint32_t f = 0; f < nfilters; f++) {
    const bool head = first_is_head && f == 0;
    if (j >= (head ? n_first : n_rest)) continue;
    const float* in = head ? src + srcchannel + j : out;
    const uint32_t in_stride = head ? nsrc : ndst;
    const AllpassSection s = b.sec[f];
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
  }
}

}  // namespace bbx

using namespace bbx;

struct bbx_allpass {
  uint32_t nch = 0, nf = 0;
  std::vector<uint32_t> delay, pos;  // pos in ring items, like RingBuffer::GetPosition()
  std::vector<float> coeff;
  std::vector<float*> ring;
  int device = 0;
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
  BBX_CUDA_TRY(cudaGetDevice(&a->device));
  a->delay.assign(delays, delays + nfilters);
  a->coeff.assign(coeffs, coeffs + nfilters);
  a->pos.assign(nfilters, 0);
  a->ring.assign(nfilters, nullptr);
  for (uint32_t f = 0; f < nfilters; f++) {
    const size_t bytes = sizeof(float) * (size_t)nchannels * delays[f];
    BBX_CUDA_TRY(cudaMalloc((void**)&a->ring[f], bytes));
    BBX_CUDA_TRY(cudaMemset(a->ring[f], 0, bytes));
  }
  *out = guard.release();
  return BBX_OK;
}

int bbx_allpass_destroy(bbx_allpass* a) {
  if (!a) return BBX_OK;
  DeviceGuard dg(a->device);
  for (float* r : a->ring) cudaFree(r);
  delete a;
  return BBX_OK;
}

int bbx_allpass_process_dev(bbx_allpass* a, const float* src, float* dst, uint32_t srcchannel, uint32_t nsrcchannels,
                            uint32_t dstchannel, uint32_t ndstchannels, uint32_t nframes, void* stream) {
  BBX_REQUIRE(a != nullptr, "bbx_allpass_process: null argument");
  if (!a->nf || !nframes) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_allpass_process: null buffer");
  // channels that fit the geometries (src/AllPassFilter.h:108-111); the single-channel branch does not clamp.  The first
  // section sees (src, dst), the following ones (dst, dst): the chain re-points src at dst after its first filter
  uint32_t n_first = a->nch, n_rest = a->nch;
  if (a->nch != 1) {
    const uint32_t fit_src = nsrcchannels >= srcchannel ? nsrcchannels - srcchannel : 0u;
    const uint32_t fit_dst = ndstchannels >= dstchannel ? ndstchannels - dstchannel : 0u;
    n_first = std::min(a->nch, std::min(fit_src, fit_dst));
    n_rest = std::min(a->nch, fit_dst);
  } else {
    BBX_REQUIRE(srcchannel < nsrcchannels && dstchannel < ndstchannels, "bbx_allpass_process: channel outside the buffers");
  }
  BBX_REQUIRE(!(src == dst && (srcchannel != dstchannel || nsrcchannels != ndstchannels)),
              "bbx_allpass_process: in-place use needs equal source and destination channels (src/AllPassFilter.h:87)");
  DeviceGuard dg(a->device);
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t n = std::max(n_first, n_rest);
  for (uint32_t f0 = 0; f0 < a->nf && n; f0 += kAllpassBatch) {
    AllpassBatch b;
    const uint32_t nb = std::min(kAllpassBatch, a->nf - f0);
    for (uint32_t k = 0; k < nb; k++) {
      const uint32_t f = f0 + k;
      b.sec[k].ring = a->ring[f];
      b.sec[k].delay = a->delay[f];
      b.sec[k].slot = a->pos[f] / a->nch;
      b.sec[k].coeff = a->coeff[f];
      b.sec[k].pad = 0;
    }
    k_allpass<<<ceil_div(n, 128u), 128, 0, st>>>(b, nb, f0 == 0 ? 1u : 0u, src, dst, a->nch, n_first, n_rest, srcchannel,
                                                nsrcchannels, dstchannel, ndstchannels, nframes);
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
  DeviceGuard dg(a->device);
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
    DeviceGuard dg(a->device);
    cudaStreamSynchronize(cudaStreamPerThread);
    cudaMemcpy(ring, a->ring[filter], sizeof(float) * n, cudaMemcpyDeviceToHost);
  }
  return a->pos[filter];
}

}  // extern "C"
