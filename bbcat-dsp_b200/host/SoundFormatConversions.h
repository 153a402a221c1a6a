// Drop-in C++ surface for bbcat-dsp's sample-format entry points, backed by libbbx (CUDA, sm_100a).
// Same names, argument meaning and error behaviour as src/SoundFormatConversions.h:20-198 of the
// reference; the work happens in bbx_transfer_samples (include/bbx.h).
//
// Ditherer (src/SoundFormatConversions.h:39-54): the class is here with the reference's three virtuals.  A NULL ditherer
// and converters without a dither call site take the plain GPU path.  TPDFDitherer (Dither_TPDF, which the reference's enum
// names and its tree never implements) runs on the device (bbx_transfer_samples_dither).  Any other subclass is honoured
// exactly as the reference would: the rectangle is loaded into the converter's intermediate type (sint32 / float / double)
// by the GPU path, the caller's Dither() runs on the host for every sample -- same arguments (the frame LOOP counter as
// `channel`, the converter's bit count) and the same call order as the reference's loops, reversed frame direction
// included -- and the GPU path quantises the result.
#pragma once

#include <stdint.h>

#include <vector>

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

class Ditherer {
public:
  Ditherer() {}
  virtual ~Ditherer() {}

  virtual void Dither(uint_t channel, sint32_t& data, uint_t bits) { (void)channel; (void)data; (void)bits; }
  virtual void Dither(uint_t channel, float& data, uint_t bits) { (void)channel; (void)data; (void)bits; }
  virtual void Dither(uint_t channel, double& data, uint_t bits) { (void)channel; (void)data; (void)bits; }
};

typedef enum {
  Dither_None = BBX_DITHER_NONE,
  Dither_TPDF = BBX_DITHER_TPDF,
} Dither_t;

// Dither_TPDF on the device: no host hook, the noise is generated inside the conversion kernel (law: include/bbx.h, a7).
// Every transfer draws from a fresh stream (the seed advances per call), so repeated blocks do not repeat their noise.
class TPDFDitherer : public Ditherer {
public:
  explicit TPDFDitherer(uint64_t _seed = 0x5eed5eedull) : seed(_seed) {}
  uint64_t NextSeed() { return seed += 0x9E3779B97F4A7C15ull; }

private:
  uint64_t seed;
};

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

namespace detail {

// host-hook path for an arbitrary Ditherer subclass on a converter that has a dither call site
template <typename T>
inline void TransferSamplesHooked(const void* vsrc, SampleFormat_t srctype, bool src_be, uint_t src_channel, uint_t src_channels,
                                  void* vdst, SampleFormat_t dsttype, bool dst_be, uint_t dst_channel, uint_t dst_channels,
                                  uint_t nchannels, uint_t nframes, Ditherer* ditherer, uint_t bits, SampleFormat_t midtype) {
  // the reference's loops run over the rectangle AFTER the sanity checks (a contiguous rectangle collapses to one frame)
  if (!BlockTransferSanityChecks(src_channel, src_channels, dst_channel, dst_channels, nchannels, nframes)) return;
  // frames run backwards when the destination frame is the longer one (src/SoundFormatConversions.cpp:178-185); the hook
  // still sees the loop counter counting up.  Channels run forwards in every converter that dithers.
  const bool reversed = (size_t)dst_channels * GetBytesPerSample(dsttype) > (size_t)src_channels * GetBytesPerSample(srctype);
  if (nframes == 1) {  // one frame (possibly a collapsed contiguous rectangle): the frame strides no longer matter, but
    src_channels = src_channel + nchannels;  // the entry points below re-run the checks and must see a consistent geometry
    dst_channels = dst_channel + nchannels;
  }
  std::vector<T> mid((size_t)nchannels * nframes);
  if (bbx_transfer_samples(vsrc, (int)srctype, src_be, src_channel, src_channels, &mid[0], (int)midtype, false, 0, nchannels,
                           nchannels, nframes) != BBX_OK)
    return;
  for (uint_t i = 0; i < nframes; i++) {
    T* frame = &mid[(size_t)(reversed ? nframes - 1 - i : i) * nchannels];
    for (uint_t j = 0; j < nchannels; j++) ditherer->Dither(i, frame[j], bits);
  }
  (void)bbx_transfer_samples(&mid[0], (int)midtype, false, 0, nchannels, vdst, (int)dsttype, dst_be, dst_channel, dst_channels,
                             nchannels, nframes);
}

}  // namespace detail

inline void TransferSamples(const void* vsrc, SampleFormat_t srctype, bool src_be, uint_t src_channel, uint_t src_channels,
                            void* vdst, SampleFormat_t dsttype, bool dst_be, uint_t dst_channel, uint_t dst_channels,
                            uint_t nchannels = ~0u, uint_t nframes = 1, Ditherer* ditherer = 0) {
  const int bits = ditherer ? bbx_dither_bits((int)srctype, (int)dsttype) : -1;
  if (bits < 0) {  // no ditherer, or a converter that never calls it
    (void)bbx_transfer_samples(vsrc, (int)srctype, src_be, src_channel, src_channels, vdst, (int)dsttype, dst_be, dst_channel,
                               dst_channels, nchannels, nframes);
  } else if (TPDFDitherer* tpdf = dynamic_cast<TPDFDitherer*>(ditherer)) {
    (void)bbx_transfer_samples_dither(vsrc, (int)srctype, src_be, src_channel, src_channels, vdst, (int)dsttype, dst_be,
                                      dst_channel, dst_channels, nchannels, nframes, BBX_DITHER_TPDF, tpdf->NextSeed());
  } else if (srctype == SampleFormat_Float) {
    detail::TransferSamplesHooked<float>(vsrc, srctype, src_be, src_channel, src_channels, vdst, dsttype, dst_be, dst_channel,
                                         dst_channels, nchannels, nframes, ditherer, (uint_t)bits, SampleFormat_Float);
  } else if (srctype == SampleFormat_Double) {
    detail::TransferSamplesHooked<double>(vsrc, srctype, src_be, src_channel, src_channels, vdst, dsttype, dst_be, dst_channel,
                                          dst_channels, nchannels, nframes, ditherer, (uint_t)bits, SampleFormat_Double);
  } else {
    detail::TransferSamplesHooked<sint32_t>(vsrc, srctype, src_be, src_channel, src_channels, vdst, dsttype, dst_be, dst_channel,
                                            dst_channels, nchannels, nframes, ditherer, (uint_t)bits, SampleFormat_32bit);
  }
}

inline void TransferSamplesLinear(const void* vsrc, SampleFormat_t srctype, void* vdst, SampleFormat_t dsttype,
                                  uint_t nsamples = 1, Ditherer* ditherer = 0) {
  // src/SoundFormatConversions.cpp:204-219: one frame of nsamples channels, machine byte order
  if ((int)srctype <= 0 || srctype >= SampleFormat_Count || (int)dsttype <= 0 || dsttype >= SampleFormat_Count || !nsamples) return;
  TransferSamples(vsrc, srctype, false, 0, nsamples, vdst, dsttype, false, 0, nsamples, nsamples, 1, ditherer);
}

template <typename T1, typename T2>
void TransferSamples(const T1* src, uint_t src_channel, uint_t src_channels, T2* dst, uint_t dst_channel, uint_t dst_channels,
                     uint_t nchannels = ~0u, uint_t nframes = 1, Ditherer* ditherer = 0) {
  TransferSamples((const void*)src, SampleFormatOf(*src), false, src_channel, src_channels, (void*)dst, SampleFormatOf(*dst),
                  false, dst_channel, dst_channels, nchannels, nframes, ditherer);
}

}  // namespace bbcat
