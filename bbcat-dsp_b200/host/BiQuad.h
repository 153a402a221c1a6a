// BiQuadCoeffs / BiQuad surface (bbcat-dsp src/BiQuad.h:27-245): coefficient object with design (CalcCoeffs),
// explicit coefficients (SetCoeffs) and ramping, shared by a bank of per-channel filters that run on the GPU.
// Thin RAII wrapper over the bbx_biquad C ABI; filter type values equal BiQuadCoeffs::Filter_t.
#pragma once

#include <stdexcept>
#include <string>

#include "SoundFormatConversions.h"

namespace bbcat {

class BiQuadBank {
public:
  typedef enum { FLAT, LPF6, HPF6, LPF12, HPF12, BPF, NOTCH, PEQ, LSH, HSH } Filter_t;  // src/BiQuad.h:31-42
  typedef struct {
    double num0, num1, num2, den1, den2;
  } COEFFS;

  explicit BiQuadBank(uint_t channels) : b(0) { Check(bbx_biquad_create(channels, &b)); }
  ~BiQuadBank() { bbx_biquad_destroy(b); }

  // BiQuadCoeffs::SetCoeffs (interpolation time in SAMPLES) / CalcCoeffs (interpolation time in SECONDS)
  void SetCoeffs(double num0, double num1 = 0.0, double num2 = 0.0, double den1 = 0.0, double den2 = 0.0, double interp_samples = 0.0) {
    const double c[5] = {num0, num1, num2, den1, den2};
    Check(bbx_biquad_set_coeffs(b, c, interp_samples));
  }
  void CalcCoeffs(Filter_t type, double freq, double fs, double gain = 0.0, double bandwidth = 1.0, double interp_time = 0.0) {
    Check(bbx_biquad_calc(b, (int)type, freq, fs, gain, bandwidth, interp_time));
  }
  COEFFS GetCurrent() const {
    double c[5];
    Check(bbx_biquad_get_state(b, 0, c, 0));
    COEFFS r = {c[0], c[1], c[2], c[3], c[4]};
    return r;
  }
  // BiQuad::Process(filters, src, dst, nchannels, nsrcchannels, ndstchannels, nframes, coeffs); dst may equal src
  void Process(const Sample_t* src, Sample_t* dst, uint_t nchannels, uint_t nsrcchannels, uint_t ndstchannels, uint_t nframes) {
    Check(bbx_biquad_process(b, src, dst, nchannels, nsrcchannels, ndstchannels, nframes));
  }
  void Reset() { Check(bbx_biquad_reset(b)); }

private:
  static void Check(int rc) {
    if (rc != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }
  BiQuadBank(const BiQuadBank&);
  BiQuadBank& operator=(const BiQuadBank&);
  bbx_biquad* b;
};

// BiQuadFilterBank surface (src/BiQuad.h:247-353): filters in series on every channel, one coefficient object per filter.
// Process runs all filters in ONE pass over the block on the GPU (k_fbank) and is bit-exact against the reference's
// filter-by-filter loop (src/BiQuad.cpp:639-662).  GetFilterCoeffs(i) returns a small handle with the reference's
// SetCoeffs / CalcCoeffs signatures instead of a BiQuadCoeffs pointer.
class BiQuadFilterBank {
public:
  typedef BiQuadBank::Filter_t Filter_t;
  typedef BiQuadBank::COEFFS COEFFS;

  class Coeffs {  // what GetFilterCoeffs(i) hands out: the coefficient object of filter i
  public:
    void SetCoeffs(double num0, double num1 = 0.0, double num2 = 0.0, double den1 = 0.0, double den2 = 0.0, double interp_samples = 0.0) {
      const double c[5] = {num0, num1, num2, den1, den2};
      Check(bbx_fbank_set_coeffs(f, i, c, interp_samples));
    }
    void CalcCoeffs(Filter_t type, double freq, double fs, double gain = 0.0, double bandwidth = 1.0, double interp_time = 0.0) {
      Check(bbx_fbank_calc(f, i, (int)type, freq, fs, gain, bandwidth, interp_time));
    }
    COEFFS GetCurrent() const {
      double c[5];
      Check(bbx_fbank_get_state(f, i, 0, c, 0));
      COEFFS r = {c[0], c[1], c[2], c[3], c[4]};
      return r;
    }
    bool Valid() const { return f != 0; }

  private:
    friend class BiQuadFilterBank;
    Coeffs(bbx_fbank* _f, uint_t _i) : f(_f), i(_i) {}
    bbx_fbank* f;
    uint_t i;
  };

  BiQuadFilterBank() : f(0) { Check(bbx_fbank_create(0, 0, &f)); }
  ~BiQuadFilterBank() { bbx_fbank_destroy(f); }

  void SetFilters(uint_t n) { Check(bbx_fbank_set_filters(f, n)); }
  void AddFilter(const COEFFS& c) {
    const double c5[5] = {c.num0, c.num1, c.num2, c.den1, c.den2};
    Check(bbx_fbank_add_filter(f, c5));
  }
  uint_t GetFilters() const {
    uint32_t n = 0;
    Check(bbx_fbank_get_size(f, 0, &n));
    return n;
  }
  void Reset() { Check(bbx_fbank_reset(f)); }
  void SetChannels(uint_t n) { Check(bbx_fbank_set_channels(f, n)); }
  uint_t GetChannels() const {
    uint32_t n = 0;
    Check(bbx_fbank_get_size(f, &n, 0));
    return n;
  }
  // the reference returns NULL beyond the last filter; here the handle's Valid() is false
  Coeffs GetFilterCoeffs(uint_t i) { return Coeffs(i < GetFilters() ? f : 0, i); }
  void Process(const Sample_t* src, Sample_t* dst, uint_t nchannels, uint_t nsrcchannels, uint_t ndstchannels, uint_t nframes) {
    Check(bbx_fbank_process(f, src, dst, nchannels, nsrcchannels, ndstchannels, nframes));
  }

private:
  static void Check(int rc) {
    if (rc != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }
  BiQuadFilterBank(const BiQuadFilterBank&);
  BiQuadFilterBank& operator=(const BiQuadFilterBank&);
  bbx_fbank* f;
};

// BiQuadCascade surface (src/BiQuad.h:373-792) for a bank of independent per-channel cascades on the GPU: numfilters (1..12)
// float biquads per channel, plain or "vectorised" (pipelined) Tick, coefficient vector (g, b1[0], b2[0], a1[0], a2[0], ...).
// The reference object filters one mono stream (ProcessCascade(input, dest, blocksize)); the bank runs `channels` of
// them in one call, interleaved or planar.
class BiQuadCascadeBank {
public:
  BiQuadCascadeBank(uint_t channels, uint_t numfilters, bool vectorise = true, bool unroll = true) : c(0), nch(channels) {
    Check(bbx_cascade_create(channels, numfilters, vectorise, unroll, &c));
  }
  ~BiQuadCascadeBank() { bbx_cascade_destroy(c); }
  // BiQuadCascade::SetCoefficients(const std::vector<float>&): false (and no change) on a wrong length, like the reference
  bool SetCoefficients(const float* coefficients, uint_t n) { return bbx_cascade_set_coefficients(c, ~0u, coefficients, n) == BBX_OK; }
  bool SetCoefficients(uint_t channel, const float* coefficients, uint_t n) {
    return bbx_cascade_set_coefficients(c, channel, coefficients, n) == BBX_OK;
  }
  void Reset() { Check(bbx_cascade_reset(c)); }
  // ProcessCascade on every channel: [blocksize][channels] when interleaved, else [channels][blocksize]
  void ProcessCascade(const float* input, float* dest, uint_t blocksize, bool interleaved = true) {
    Check(bbx_cascade_process(c, input, dest, blocksize, interleaved));
  }

private:
  static void Check(int rc) {
    if (rc != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }
  BiQuadCascadeBank(const BiQuadCascadeBank&);
  BiQuadCascadeBank& operator=(const BiQuadCascadeBank&);
  bbx_cascade* c;
  uint_t nch;
};

}  // namespace bbcat
