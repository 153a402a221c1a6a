// AllPassFilterChain<float> surface (bbcat-dsp src/AllPassFilter.h:130-262): a chain of Schroeder all-pass sections
// over interleaved channels, state (the per-section rings) in HBM.  Thin RAII wrapper over the bbx_allpass C ABI.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "SoundFormatConversions.h"

namespace bbcat {

class AllPassFilterChain {
public:
  // delays in frames (>= 1) and coefficients per section, like AllPassFilterChain(nchannels, nfilters, delays, coeffs)
  AllPassFilterChain(uint_t nchannels, uint_t nfilters, const uint_t* delays, const float* coeffs) : a(0) {
    Check(bbx_allpass_create(nchannels, nfilters, delays, coeffs, &a));
  }
  ~AllPassFilterChain() { bbx_allpass_destroy(a); }
  void Process(const float* src, float* dst, uint_t srcchannel, uint_t nsrcchannels, uint_t dstchannel, uint_t ndstchannels,
               uint_t nframes = 1) {
    Check(bbx_allpass_process(a, src, dst, srcchannel, nsrcchannels, dstchannel, ndstchannels, nframes));
  }

private:
  static void Check(int rc) {
    if (rc != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }
  AllPassFilterChain(const AllPassFilterChain&);
  AllPassFilterChain& operator=(const AllPassFilterChain&);
  bbx_allpass* a;
};

}  // namespace bbcat
