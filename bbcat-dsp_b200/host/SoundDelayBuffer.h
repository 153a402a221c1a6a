// SoundDelayBuffer with the reference's method set (src/SoundDelayBuffer.h:23-95); the ring lives in HBM.
#pragma once

#include "SoundFormatConversions.h"

namespace bbcat {

class SoundDelayBuffer {
public:
  SoundDelayBuffer() : h(0) { bbx_delay_create(&h); }
  virtual ~SoundDelayBuffer() { bbx_delay_destroy(h); }

  virtual void SetSize(uint_t chans, uint_t length, SampleFormat_t type = SampleFormat_Float) {
    bbx_delay_set_size(h, chans, length, (int)type);
  }
  uint_t GetChannels() const { return bbx_delay_get_channels(h); }
  uint_t GetLength() const { return bbx_delay_get_length(h); }
  uint_t GetWritePosition() const { return bbx_delay_get_write_position(h); }
  SampleFormat_t GetFormat() const { return (SampleFormat_t)bbx_delay_get_format(h); }
  // GetBuffer(): the ring is device memory; this is a DEVICE pointer (for bbx_*_dev entry points)
  const void* GetDeviceBuffer() const { return bbx_delay_get_buffer_dev(h); }

  virtual uint_t WriteSamples(const uint8_t* src, SampleFormat_t srcformat, uint_t channel = 0, uint_t nchannels = ~0u,
                              uint_t nframes = 1) {
    return bbx_delay_write_samples(h, src, (int)srcformat, channel, nchannels, nframes);
  }
  template <typename T>
  uint_t WriteSamples(const T* src, uint_t channel = 0, uint_t nchannels = ~0u, uint_t nframes = 1) {
    return WriteSamples((const uint8_t*)src, SampleFormatOf(src), channel, nchannels, nframes);
  }
  virtual void IncrementWritePosition(uint_t nframes = 1) { bbx_delay_increment_write_position(h, nframes); }
  virtual uint_t ReadSamples(uint8_t* dst, SampleFormat_t dstformat, uint_t delay, uint_t channel = 0, uint_t nchannels = ~0u,
                             uint_t nframes = 1) {
    return bbx_delay_read_samples(h, dst, (int)dstformat, delay, channel, nchannels, nframes);
  }
  template <typename T>
  uint_t ReadSamples(T* dst, uint_t delay, uint_t channel = 0, uint_t nchannels = ~0u, uint_t nframes = 1) {
    return ReadSamples((uint8_t*)dst, SampleFormatOf(dst), delay, channel, nchannels, nframes);
  }
  virtual Sample_t ReadSample(uint_t channel, uint_t delay) const { return bbx_delay_read_sample(h, channel, delay); }

protected:
  struct RingTag {};
  explicit SoundDelayBuffer(RingTag) : h(0) { bbx_ring_create(&h); }
  bbx_delay* h;

private:
  SoundDelayBuffer(const SoundDelayBuffer&);
  SoundDelayBuffer& operator=(const SoundDelayBuffer&);
};

// SoundRingBuffer (src/SoundDelayBuffer.h:105-181): the overrides of SetSize / WriteSamples / IncrementWritePosition /
// ReadSamples live behind the same C entry points (a handle made by bbx_ring_create applies the read-position limits)
class SoundRingBuffer : public SoundDelayBuffer {
public:
  SoundRingBuffer() : SoundDelayBuffer(RingTag()) {}
  virtual uint_t GetReadPosition() const { return bbx_ring_get_read_position(h); }
  virtual uint_t GetReadFramesAvailable() const { return bbx_ring_get_read_frames_available(h); }
  virtual uint_t GetWriteFramesAvailable() const { return bbx_ring_get_write_frames_available(h); }
  virtual void IncrementReadPosition(uint_t nframes = 1) { bbx_ring_increment_read_position(h, nframes); }
};

}  // namespace bbcat
