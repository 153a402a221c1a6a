/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of MultilayerBuffer<float> (MultilayerBuffer.h:19-431): an output collection bus that
 * several renderers ("layers") with different block sizes mix into; frames become readable once
 * every layer has written them.
 *   WriteLayer   ReserveSpace + MixSamples at the layer's position + LayerWritten   (.h:185-202)
 *   LayerWritten positions[layer] += n; minposition = min over layers; maxposition   (.h:227-250)
 *   ReadBuffer   n = min(n, minposition); TransferSamples or MixSamples; BufferRead  (.h:281-308)
 *   BufferRead   shift the unread frames to the front, zero the freed tail           (.h:383-407)
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

struct orc_mlb {
  float* buffer;
  size_t size;      /* floats in use (vector size) */
  size_t capacity;  /* floats allocated */
  unsigned* positions;
  unsigned layers, channels, minposition, maxposition;
};

orc_mlb* orc_mlb_create(unsigned channels, unsigned layers) {
  orc_mlb* m = (orc_mlb*)calloc(1, sizeof(*m));
  m->channels = channels;
  m->layers = layers;
  m->positions = (unsigned*)calloc(layers ? layers : 1, sizeof(unsigned));
  return m;
}

void orc_mlb_destroy(orc_mlb* m) {
  if (!m) return;
  free(m->buffer);
  free(m->positions);
  free(m);
}

static void reserve_space(orc_mlb* m, unsigned layer, unsigned nframes) {
  if (layer >= m->layers) return;
  size_t need = (size_t)(m->positions[layer] + nframes) * m->channels;
  if (need > m->size) { /* std::vector::resize: new elements are zero */
    if (need > m->capacity) {
      size_t cap = m->capacity ? m->capacity : 1024;
      while (cap < need) cap *= 2;
      m->buffer = (float*)realloc(m->buffer, cap * sizeof(float));
      m->capacity = cap;
    }
    memset(m->buffer + m->size, 0, (need - m->size) * sizeof(float));
    m->size = need;
  }
}

static unsigned layer_written(orc_mlb* m, unsigned layer, unsigned nframes) {
  unsigned i;
  if (layer < m->layers) {
    reserve_space(m, layer, nframes);
    m->positions[layer] += nframes;
    for (i = 0; i < m->layers; i++) {
      if (i == 0) m->minposition = m->positions[i];
      else if (m->positions[i] < m->minposition) m->minposition = m->positions[i];
    }
    if (m->positions[layer] > m->maxposition) m->maxposition = m->positions[layer];
  }
  return m->minposition;
}

void orc_mlb_write_layer(orc_mlb* m, unsigned layer, const float* src, unsigned srcchannel, unsigned nsrcchannels,
                         unsigned dstchannel, unsigned nchannels, unsigned nframes) {
  if (layer >= m->layers) return;
  reserve_space(m, layer, nframes);
  orc_mix_samples_f32(src, srcchannel, nsrcchannels, m->buffer + (size_t)m->positions[layer] * m->channels, dstchannel,
                      m->channels, nchannels, nframes, 1.0f);
  layer_written(m, layer, nframes);
}

unsigned orc_mlb_available_frames(const orc_mlb* m) { return m->minposition; }

static unsigned buffer_read(orc_mlb* m, unsigned nframes) {
  unsigned i;
  if (nframes > m->minposition) nframes = m->minposition;
  if (nframes > 0) {
    m->minposition -= nframes;
    m->maxposition -= nframes;
    for (i = 0; i < m->layers; i++) m->positions[i] -= nframes;
    if (m->maxposition)
      memmove(m->buffer, m->buffer + (size_t)nframes * m->channels, (size_t)m->maxposition * m->channels * sizeof(float));
    memset(m->buffer + (size_t)m->maxposition * m->channels, 0, (size_t)nframes * m->channels * sizeof(float));
  }
  return nframes;
}

unsigned orc_mlb_read_buffer(orc_mlb* m, unsigned srcchannel, float* dst, unsigned dstchannel, unsigned ndstchannels,
                             unsigned nchannels, unsigned nframes, int overwrite) {
  if (nframes > m->minposition) nframes = m->minposition;
  if (nframes > 0) {
    if (overwrite)
      orc_transfer_samples(m->buffer, ORC_FMT_FLOAT, 0, srcchannel, m->channels, dst, ORC_FMT_FLOAT, 0, dstchannel,
                           ndstchannels, nchannels, nframes);
    else
      orc_mix_samples_f32(m->buffer, srcchannel, m->channels, dst, dstchannel, ndstchannels, nchannels, nframes, 1.0f);
    buffer_read(m, nframes);
  }
  return nframes;
}
