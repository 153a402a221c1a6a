/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of SoundDelayBuffer (SoundDelayBuffer.cpp:11-170, .h:33-67): an
 * interleaved circular buffer [buflen][channels] in any sample format with
 * wrap-split, converting writes and delayed reads.  The write position only moves
 * through IncrementWritePosition().
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

struct orc_delay {
  uint8_t* buf;
  int format;
  unsigned channels, bytesperframe, buflen, writepos;
  /* SoundRingBuffer (SoundDelayBuffer.h:105-181, .cpp:195-304): the same buffer with a read position that limits
   * writes, reads and both increments; the entry points below branch where the reference dispatches virtually */
  int is_ring;
  unsigned readpos;
};

static unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

orc_delay* orc_delay_create(void) {
  orc_delay* d = (orc_delay*)calloc(1, sizeof(*d));
  d->format = ORC_FMT_FLOAT;
  return d;
}

orc_delay* orc_ring_create(void) {
  orc_delay* d = orc_delay_create();
  d->is_ring = 1;
  return d;
}

unsigned orc_ring_get_read_position(const orc_delay* d) { return d->readpos; }
/* SoundDelayBuffer.h:122-123 (the write side keeps one frame free so that the positions never coincide after a write) */
unsigned orc_ring_get_read_frames_available(const orc_delay* d) {
  return d->buflen ? (d->writepos + d->buflen - d->readpos) % d->buflen : 0;
}
unsigned orc_ring_get_write_frames_available(const orc_delay* d) {
  return d->buflen ? (d->readpos + d->buflen - d->writepos - 1) % d->buflen : 0;
}
/* SoundDelayBuffer.h:175 */
void orc_ring_increment_read_position(orc_delay* d, unsigned nframes) {
  if (!d->buflen) return;
  nframes = nframes < orc_ring_get_read_frames_available(d) ? nframes : orc_ring_get_read_frames_available(d);
  d->readpos = (d->readpos + nframes) % d->buflen;
}

void orc_delay_destroy(orc_delay* d) {
  if (!d) return;
  free(d->buf);
  free(d);
}

void orc_delay_set_size(orc_delay* d, unsigned chans, unsigned length, int fmt) {
  chans = umax(chans, 1u);
  length = umax(length, 1u);
  if (chans == d->channels && length == d->buflen && fmt == d->format) return;
  unsigned bps = orc_get_bytes_per_sample(fmt);
  uint8_t* nb = (uint8_t*)calloc((size_t)chans * length * bps, 1);
  if (!nb) return;
  if (d->buf) {
    /* SoundDelayBuffer.cpp:46-49 maps the old contents across with the OLD format tag on both sides
     * and the OLD frame count.  That is only memory-safe when the format is unchanged and the buffer
     * does not shrink; outside that domain the reference overruns its new allocation (UB), so here
     * the copy is clamped to the new length and skipped on a format change (contents dropped). */
    if (fmt == d->format)
      orc_transfer_samples(d->buf, d->format, 0, 0, d->channels, nb, d->format, 0, 0, chans, ~0u,
                           umin(d->buflen, length));
    free(d->buf);
  }
  d->buf = nb;
  d->channels = chans;
  d->buflen = length;
  d->format = fmt;
  d->writepos %= d->buflen;
  d->bytesperframe = chans * bps;
  if (d->is_ring) d->readpos %= d->buflen; /* .cpp:214-218 (a no-op on the unchanged-geometry early return above) */
}

unsigned orc_delay_get_channels(const orc_delay* d) { return d->channels; }
unsigned orc_delay_get_length(const orc_delay* d) { return d->buflen; }
unsigned orc_delay_get_write_position(const orc_delay* d) { return d->writepos; }
int orc_delay_get_format(const orc_delay* d) { return d->format; }

unsigned orc_delay_write_samples(orc_delay* d, const void* vsrc, int srcformat, unsigned channel, unsigned nchannels,
                                 unsigned nframes) {
  unsigned frames = 0;
  if (!d->buf) return 0;
  if (d->is_ring) nframes = umin(nframes, orc_ring_get_write_frames_available(d)); /* .cpp:234-254 */
  const uint8_t* src = (const uint8_t*)vsrc;
  unsigned srclen = orc_get_bytes_per_sample(srcformat), pos = d->writepos;
  channel = umin(channel, d->channels - 1);
  nchannels = umin(nchannels, d->channels - channel);
  while (nframes) {
    uint8_t* dst = d->buf + (size_t)pos * d->bytesperframe;
    unsigned n = umin(nframes, d->buflen - pos);
    orc_transfer_samples(src, srcformat, 0, 0, nchannels, dst, d->format, 0, channel, d->channels, nchannels, n);
    src += (size_t)nchannels * srclen * n;
    pos = (pos + n) % d->buflen;
    nframes -= n;
    frames += n;
  }
  return frames;
}

void orc_delay_increment_write_position(orc_delay* d, unsigned nframes) {
  if (d->is_ring && d->buflen) nframes = umin(nframes, orc_ring_get_write_frames_available(d)); /* .h:148 */
  if (d->buflen) d->writepos = (d->writepos + nframes) % d->buflen;
}

unsigned orc_delay_read_samples(orc_delay* d, void* vdst, int dstformat, unsigned delay, unsigned channel,
                                unsigned nchannels, unsigned nframes) {
  unsigned frames = 0;
  if (!d->buf) return 0;
  if (d->is_ring) {
    /* .cpp:274-303: the delay is limited to (read - write) mod length, the frame count to (write + delay - read) mod
     * length, then the base class reads relative to the WRITE position */
    delay = umin(delay, (d->readpos + d->buflen - d->writepos) % d->buflen);
    nframes = umin(nframes, (d->writepos + d->buflen + delay - d->readpos) % d->buflen);
  }
  uint8_t* dst = (uint8_t*)vdst;
  unsigned dstlen = orc_get_bytes_per_sample(dstformat);
  unsigned pos = (d->writepos + d->buflen - delay) % d->buflen;
  channel = umin(channel, d->channels - 1);
  nchannels = umin(nchannels, d->channels - channel);
  nframes = umin(nframes, delay); /* cannot read past the write position (.cpp:147-149) */
  while (nframes) {
    const uint8_t* src = d->buf + (size_t)pos * d->bytesperframe;
    unsigned n = umin(nframes, d->buflen - pos);
    orc_transfer_samples(src, d->format, 0, channel, d->channels, dst, dstformat, 0, 0, nchannels, nchannels, n);
    dst += (size_t)nchannels * dstlen * n;
    pos = (pos + n) % d->buflen;
    nframes -= n;
    frames += n;
  }
  return frames;
}

unsigned orc_delay_copy_buffer(const orc_delay* d, void* dst, unsigned maxbytes) {
  unsigned bytes = d->bytesperframe * d->buflen;
  if (!d->buf || bytes > maxbytes) return 0;
  memcpy(dst, d->buf, bytes);
  return bytes;
}
