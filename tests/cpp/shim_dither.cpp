// TEST INFRASTRUCTURE: extern "C" door into the C++ host shim's TransferSamples with a Ditherer object, same signature as
// ref_transfer_samples_ditherer (oracle/ref_wrap.cpp) so that tests/test_dither.py can run the reference and the shim side
// by side.  mode 0 = the no-op base class, 1 = the stateful test subclass, 2 = TPDFDitherer(seed 77) on the device.
#include "SoundFormatConversions.h"
#include "test_ditherer.h"

using namespace bbcat;

extern "C" unsigned shim_transfer_samples_ditherer(const void* src, int srctype, int src_be, unsigned src_channel,
                                                   unsigned src_channels, void* dst, int dsttype, int dst_be,
                                                   unsigned dst_channel, unsigned dst_channels, unsigned nchannels,
                                                   unsigned nframes, int mode) {
  Ditherer base;
  TestDitherer test;
  TPDFDitherer tpdf(77);
  Ditherer* d = mode == 0 ? &base : (mode == 1 ? (Ditherer*)&test : (Ditherer*)&tpdf);
  TransferSamples(src, (SampleFormat_t)srctype, src_be != 0, src_channel, src_channels, dst, (SampleFormat_t)dsttype, dst_be != 0,
                  dst_channel, dst_channels, nchannels, nframes, d);
  return test.calls;
}

extern "C" unsigned long long shim_tpdf_first_seed(void) {
  TPDFDitherer tpdf(77);
  return tpdf.NextSeed();
}
