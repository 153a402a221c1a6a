// multilayer.cu -- MultilayerBuffer<float> with the bus in HBM (the first "next" row).
//
// Reference: src/MultilayerBuffer.h:19-431.  Renderers with different block sizes ("layers") mix their blocks
// into one output bus at their own write positions; frames become readable once every layer has written them.
// The host logic (positions, min/max, reserve, shift) follows the reference; the sample work runs through the
// device entry points bbx_mix_samples_f32_dev / bbx_transfer_samples_dev, so results are bit-exact.
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace bbx;

struct bbx_mlb {
  float* buf = nullptr;  // device [frames][channels]
  size_t size = 0;       // floats in use (the reference's vector size)
  size_t capacity = 0;   // floats allocated
  std::vector<uint32_t> positions;
  uint32_t channels = 0, minposition = 0, maxposition = 0;
  int device = 0;
};

namespace {

int reserve_space(bbx_mlb* m, uint32_t layer, uint32_t nframes, cudaStream_t st) {
  if (layer >= m->positions.size()) return BBX_OK;
  size_t need = (size_t)(m->positions[layer] + nframes) * m->channels;
  if (need <= m->size) return BBX_OK;
  if (need > m->capacity) {
    size_t cap = m->capacity ? m->capacity : 4096;
    while (cap < need) cap *= 2;
    float* nb = nullptr;
    BBX_CUDA_TRY(cudaMalloc((void**)&nb, cap * sizeof(float)));
    if (m->size) BBX_CUDA_TRY(cudaMemcpyAsync(nb, m->buf, m->size * sizeof(float), cudaMemcpyDeviceToDevice, st));
    BBX_CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(m->buf);
    m->buf = nb;
    m->capacity = cap;
  }
  // std::vector::resize zero-fills the new elements (.h:160-167)
  BBX_CUDA_TRY(cudaMemsetAsync(m->buf + m->size, 0, (need - m->size) * sizeof(float), st));
  m->size = need;
  return BBX_OK;
}

void layer_written(bbx_mlb* m, uint32_t layer, uint32_t nframes) {
  m->positions[layer] += nframes;
  for (size_t i = 0; i < m->positions.size(); i++)
    m->minposition = (i == 0) ? m->positions[i] : std::min(m->minposition, m->positions[i]);
  m->maxposition = std::max(m->maxposition, m->positions[layer]);
}

int buffer_read(bbx_mlb* m, uint32_t nframes, cudaStream_t st) {
  nframes = std::min(nframes, m->minposition);
  if (!nframes) return BBX_OK;
  m->minposition -= nframes;
  m->maxposition -= nframes;
  for (auto& p : m->positions) p -= nframes;
  size_t keep = (size_t)m->maxposition * m->channels, drop = (size_t)nframes * m->channels;
  if (keep) {  // overlapping move: go through scratch
    DeviceScratch& s = scratch(3);
    int rc = s.ensure(keep * sizeof(float));
    if (rc) return rc;
    BBX_CUDA_TRY(cudaMemcpyAsync(s.ptr, m->buf + drop, keep * sizeof(float), cudaMemcpyDeviceToDevice, st));
    BBX_CUDA_TRY(cudaMemcpyAsync(m->buf, s.ptr, keep * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  BBX_CUDA_TRY(cudaMemsetAsync(m->buf + keep, 0, drop * sizeof(float), st));
  return BBX_OK;
}

}  // namespace

extern "C" {

int bbx_mlb_create(uint32_t channels, uint32_t layers, bbx_mlb** out) {
  BBX_REQUIRE(out != nullptr, "bbx_mlb_create: null out pointer");
  int rc = require_device();
  if (rc) return rc;
  bbx_mlb* m = new bbx_mlb();
  if (cudaGetDevice(&m->device) != cudaSuccess) m->device = 0;
  m->channels = channels;
  m->positions.assign(layers, 0u);
  *out = m;
  return BBX_OK;
}

int bbx_mlb_destroy(bbx_mlb* m) {
  if (!m) return BBX_OK;
  DeviceGuard dg(m->device);
  cudaFree(m->buf);
  delete m;
  return BBX_OK;
}

uint32_t bbx_mlb_get_channels(const bbx_mlb* m) { return m ? m->channels : 0; }
uint32_t bbx_mlb_get_layers(const bbx_mlb* m) { return m ? (uint32_t)m->positions.size() : 0; }
uint32_t bbx_mlb_get_available_frames(const bbx_mlb* m) { return m ? m->minposition : 0; }

int bbx_mlb_write_layer(bbx_mlb* m, uint32_t layer, const float* src, uint32_t srcchannel, uint32_t nsrcchannels,
                        uint32_t dstchannel, uint32_t nchannels, uint32_t nframes) {
  BBX_REQUIRE(m && src, "bbx_mlb_write_layer: null argument");
  if (layer >= m->positions.size()) return BBX_OK;  // silent, like the reference (.h:187)
  DeviceGuard dg(m->device);
  cudaStream_t st = cudaStreamPerThread;
  int rc = reserve_space(m, layer, nframes, st);
  if (rc) return rc;
  if (nframes && nsrcchannels && m->channels) {
    DeviceScratch& s = scratch(0);
    size_t bytes = (size_t)nframes * nsrcchannels * sizeof(float);
    if ((rc = s.ensure(bytes))) return rc;
    BBX_CUDA_TRY(cudaMemcpyAsync(s.ptr, src, bytes, cudaMemcpyHostToDevice, st));
    rc = bbx_mix_samples_f32_dev((const float*)s.ptr, srcchannel, nsrcchannels, m->buf + (size_t)m->positions[layer] * m->channels,
                                 dstchannel, m->channels, nchannels, nframes, 1.0f, st);
    if (rc) return rc;
  }
  layer_written(m, layer, nframes);
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  return BBX_OK;
}

uint32_t bbx_mlb_read_buffer(bbx_mlb* m, uint32_t srcchannel, float* dst, uint32_t dstchannel, uint32_t ndstchannels,
                             uint32_t nchannels, uint32_t nframes, int overwrite) {
  if (!m || !dst) return 0;
  nframes = std::min(nframes, m->minposition);
  if (!nframes) return 0;
  DeviceGuard dg(m->device);
  cudaStream_t st = cudaStreamPerThread;
  if (ndstchannels && m->channels) {
    DeviceScratch& s = scratch(1);
    size_t bytes = (size_t)nframes * ndstchannels * sizeof(float);
    if (s.ensure(bytes)) return 0;
    // the destination rectangle keeps the caller's samples outside the written channels (and is the
    // accumulation target when mixing): round-trip it
    if (cudaMemcpyAsync(s.ptr, dst, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return 0;
    int rc = overwrite ? bbx_transfer_samples_dev(m->buf, BBX_FMT_FLOAT, 0, srcchannel, m->channels, s.ptr, BBX_FMT_FLOAT, 0,
                                                  dstchannel, ndstchannels, nchannels, nframes, st)
                       : bbx_mix_samples_f32_dev(m->buf, srcchannel, m->channels, (float*)s.ptr, dstchannel, ndstchannels,
                                                 nchannels, nframes, 1.0f, st);
    if (rc) return 0;
    if (cudaMemcpyAsync(dst, s.ptr, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 0;
  }
  if (buffer_read(m, nframes, st)) return 0;
  if (cudaStreamSynchronize(st) != cudaSuccess) return 0;
  return nframes;
}

}  // extern "C"
