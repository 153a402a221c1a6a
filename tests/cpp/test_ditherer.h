// TEST INFRASTRUCTURE: one deterministic, STATEFUL Ditherer subclass shared by the reference wrapper (oracle/ref_wrap.cpp,
// on top of the reference's own class) and the shim test library (tests/cpp/shim_dither.cpp, on top of
// bbcat-dsp_b200/host/SoundFormatConversions.h).  The perturbation depends on the call counter, the `channel` argument
// (the reference passes its frame LOOP counter there) and the bit count, so a comparison of outputs checks the hook's
// position, its arguments and the order of the calls.  Include after the header that declares bbcat::Ditherer.
#pragma once

#include <stdint.h>

namespace bbcat {

class TestDitherer : public Ditherer {
public:
  TestDitherer() : calls(0) {}
  virtual void Dither(uint_t channel, sint32_t& data, uint_t bits) {
    const uint32_t mask = bits ? ((1u << bits) - 1u) : 0u;
    const int64_t v = (int64_t)data + (int64_t)(Pattern(channel) & mask) - (int64_t)(mask >> 1);
    data = (sint32_t)(v > 2147483647ll ? 2147483647ll : (v < -2147483648ll ? -2147483648ll : v));
  }
  virtual void Dither(uint_t channel, float& data, uint_t bits) { data += (float)Noise(channel, bits); }
  virtual void Dither(uint_t channel, double& data, uint_t bits) { data += Noise(channel, bits); }
  uint32_t calls;

private:
  uint32_t Pattern(uint_t channel) { return (calls++ * 40503u + channel * 7919u) * 2654435761u >> 8; }
  double Noise(uint_t channel, uint_t bits) {
    // up to +-1 LSB of the destination word, LSB = 2^(bits - 31) of full scale
    return ((double)(Pattern(channel) & 0xffffu) / 32768.0 - 1.0) * (double)(1u << bits) / 2147483648.0;
  }
};

}  // namespace bbcat
