// Drop-in C++ surface for bbcat-dsp's sample-format entry points, backed by libbbx (CUDA, sm_100a).
// Same names, argument meaning and error behaviour as src/SoundFormatConversions.h:20-198 of the
// reference; the work happens in bbx_transfer_samples (include/bbx.h).  Ditherer must be NULL
// (the reference tree only ships the no-op base class).
#pragma once

#include <stdint.h>

#include "../../include/bbx.h"

namespace bbcat {

typedef unsigned int uint_t;
typedef int32_t sint32_t;
typedef int16_t sint16_t;
typedef float Sample_t;

typedef enum {
  SampleFormat_Unknown = BBX_FMT_UNKNOWN,
  SampleFormat_16bit = BBX_FMT_16BIT,
  SampleFormat_24bit = BBX_FMT_24BIT,
  SampleFormat_32bit = BBX_FMT_32BIT,
  SampleFormat_Float = BBX_FMT_FLOAT,
  SampleFormat_Double = BBX_FMT_DOUBLE,
  SampleFormat_Count = BBX_FMT_COUNT,
} SampleFormat_t;

class Ditherer;  // only NULL is accepted

inline SampleFormat_t SampleFormatOf(sint16_t) { return SampleFormat_16bit; }
inline SampleFormat_t SampleFormatOf(sint32_t) { return SampleFormat_32bit; }
inline SampleFormat_t SampleFormatOf(float) { return SampleFormat_Float; }
inline SampleFormat_t SampleFormatOf(double) { return SampleFormat_Double; }
inline SampleFormat_t SampleFormatOf(const sint16_t*) { return SampleFormat_16bit; }
inline SampleFormat_t SampleFormatOf(const sint32_t*) { return SampleFormat_32bit; }
inline SampleFormat_t SampleFormatOf(const float*) { return SampleFormat_Float; }
inline SampleFormat_t SampleFormatOf(const double*) { return SampleFormat_Double; }

inline uint8_t GetBitsPerSample(SampleFormat_t type) { return bbx_get_bits_per_sample((int)type); }
inline uint8_t GetBytesPerSample(SampleFormat_t type) { return bbx_get_bytes_per_sample((int)type); }

inline bool BlockTransferSanityChecks(uint_t& src_channel, uint_t& src_channels, uint_t& dst_channel, uint_t& dst_channels,
                                      uint_t& nchannels, uint_t& nframes, bool allowsinglechannel = true) {
  return bbx_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes,
                                          allowsinglechannel ? 1 : 0) != 0;
}

inline void TransferSamples(const void* vsrc, SampleFormat_t srctype, bool src_be, uint_t src_channel, uint_t src_channels,
                            void* vdst, SampleFormat_t dsttype, bool dst_be, uint_t dst_channel, uint_t dst_channels,
                            uint_t nchannels = ~0u, uint_t nframes = 1, Ditherer* ditherer = 0) {
  if (ditherer) return;  // unsupported: the GPU path has no dither hook
  (void)bbx_transfer_samples(vsrc, (int)srctype, src_be, src_channel, src_channels, vdst, (int)dsttype, dst_be, dst_channel,
                             dst_channels, nchannels, nframes);
}

inline void TransferSamplesLinear(const void* vsrc, SampleFormat_t srctype, void* vdst, SampleFormat_t dsttype,
                                  uint_t nsamples = 1, Ditherer* ditherer = 0) {
  if (ditherer) return;
  (void)bbx_transfer_samples_linear(vsrc, (int)srctype, vdst, (int)dsttype, nsamples);
}

template <typename T1, typename T2>
void TransferSamples(const T1* src, uint_t src_channel, uint_t src_channels, T2* dst, uint_t dst_channel, uint_t dst_channels,
                     uint_t nchannels = ~0u, uint_t nframes = 1, Ditherer* ditherer = 0) {
  TransferSamples((const void*)src, SampleFormatOf(*src), false, src_channel, src_channels, (void*)dst, SampleFormatOf(*dst),
                  false, dst_channel, dst_channels, nchannels, nframes, ditherer);
}

}  // namespace bbcat
