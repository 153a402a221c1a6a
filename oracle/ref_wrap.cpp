/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * extern "C" handles over the reference's own in-tree code so that tests (via
 * ctypes) can run the real bbcat-dsp functions.  This file contains no
 * reference code: it only #includes the headers where they lie under
 * /root/reference/src and forwards calls.  Built by oracle/Makefile into
 * oracle/_ref/libbbcref.so together with the unmodified reference sources
 *   SoundFormatConversions.cpp, SoundFormatRawConversions.cpp, SoundMixing.cpp,
 *   FractionalSample.cpp, SoundDelayBuffer.cpp, BiQuad.cpp
 */
#include "SoundFormatConversions.h"
#include "SoundMixing.h"
#include "Interpolator.h"
#include "FractionalSample.h"
#include "SoundDelayBuffer.h"
#include "MultilayerBuffer.h"
#include "BiQuad.h"
#include "AllPassFilter.h"

#include <vector>

#include "../tests/cpp/test_ditherer.h"

using namespace bbcat;

extern "C" {

unsigned ref_get_bits_per_sample(int fmt) { return GetBitsPerSample((SampleFormat_t)fmt); }
unsigned ref_get_bytes_per_sample(int fmt) { return GetBytesPerSample((SampleFormat_t)fmt); }

int ref_block_transfer_sanity_checks(unsigned* src_channel, unsigned* src_channels, unsigned* dst_channel,
                                     unsigned* dst_channels, unsigned* nchannels, unsigned* nframes,
                                     int allowsinglechannel) {
  return BlockTransferSanityChecks(*src_channel, *src_channels, *dst_channel, *dst_channels, *nchannels, *nframes,
                                   allowsinglechannel != 0)
             ? 1
             : 0;
}

void ref_transfer_samples(const void* src, int srctype, int src_be, unsigned src_channel, unsigned src_channels,
                          void* dst, int dsttype, int dst_be, unsigned dst_channel, unsigned dst_channels,
                          unsigned nchannels, unsigned nframes) {
  TransferSamples(src, (SampleFormat_t)srctype, src_be != 0, src_channel, src_channels, dst, (SampleFormat_t)dsttype,
                  dst_be != 0, dst_channel, dst_channels, nchannels, nframes, NULL);
}

/* the reference's TransferSamples with a Ditherer: mode 0 = its own no-op base class, 1 = the stateful test subclass
 * (tests/cpp/test_ditherer.h); returns the number of hook calls the test subclass saw */
unsigned ref_transfer_samples_ditherer(const void* src, int srctype, int src_be, unsigned src_channel, unsigned src_channels,
                                       void* dst, int dsttype, int dst_be, unsigned dst_channel, unsigned dst_channels,
                                       unsigned nchannels, unsigned nframes, int mode) {
  Ditherer base;
  TestDitherer test;
  TransferSamples(src, (SampleFormat_t)srctype, src_be != 0, src_channel, src_channels, dst, (SampleFormat_t)dsttype,
                  dst_be != 0, dst_channel, dst_channels, nchannels, nframes, mode ? (Ditherer*)&test : &base);
  return test.calls;
}

void ref_transfer_samples_linear(const void* src, int srctype, void* dst, int dsttype, unsigned nsamples) {
  TransferSamplesLinear(src, (SampleFormat_t)srctype, dst, (SampleFormat_t)dsttype, nsamples, NULL);
}

void ref_mix_samples_f32(const float* src, unsigned src_channel, unsigned src_channels, float* dst,
                         unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes, float mul) {
  MixSamples<float>(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul);
}

void ref_mix_samples_f64(const double* src, unsigned src_channel, unsigned src_channels, double* dst,
                         unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes,
                         double mul) {
  MixSamples<double>(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul);
}

/* interp_state = {target, current}; updated in place like the caller's Interpolator object */
void ref_mix_samples_interp(const float* src, unsigned src_channel, unsigned src_channels, float* dst,
                            unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes,
                            float* interp_state, float inc) {
  Interpolator interp(interp_state[0], interp_state[1]);
  MixSamples(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, interp, inc);
  interp_state[0] = interp.GetTarget();
  interp_state[1] = (float)interp;
}

void ref_interpolator_step(float* interp_state, float inc, unsigned nsteps) {
  Interpolator interp(interp_state[0], interp_state[1]);
  for (unsigned i = 0; i < nsteps; i++) interp += inc;
  interp_state[0] = interp.GetTarget();
  interp_state[1] = (float)interp;
}

unsigned ref_fractional_sample_additional_delay_required(void) { return FractionalSampleAdditionalDelayRequired(); }

void ref_fractional_samples_f32(const float* buffer, unsigned channel, unsigned channels, unsigned length,
                                const double* pos, unsigned n, double* out) {
  for (unsigned i = 0; i < n; i++) out[i] = FractionalSample(buffer, channel, channels, length, pos[i]);
}

void ref_fractional_samples_f64(const double* buffer, unsigned channel, unsigned channels, unsigned length,
                                const double* pos, unsigned n, double* out) {
  for (unsigned i = 0; i < n; i++) out[i] = FractionalSample(buffer, channel, channels, length, pos[i]);
}

/* ---- SoundDelayBuffer ---- */
void* ref_delay_create(void) { return new SoundDelayBuffer(); }
/* SoundRingBuffer through the same entry points (its overrides are virtual) plus the read-position methods */
void* ref_ring_create(void) { return static_cast<SoundDelayBuffer*>(new SoundRingBuffer()); }
unsigned ref_ring_get_read_position(void* h) { return static_cast<SoundRingBuffer*>((SoundDelayBuffer*)h)->GetReadPosition(); }
unsigned ref_ring_get_read_frames_available(void* h) { return static_cast<SoundRingBuffer*>((SoundDelayBuffer*)h)->GetReadFramesAvailable(); }
unsigned ref_ring_get_write_frames_available(void* h) { return static_cast<SoundRingBuffer*>((SoundDelayBuffer*)h)->GetWriteFramesAvailable(); }
void ref_ring_increment_read_position(void* h, unsigned nframes) { static_cast<SoundRingBuffer*>((SoundDelayBuffer*)h)->IncrementReadPosition(nframes); }
void ref_delay_destroy(void* h) { delete (SoundDelayBuffer*)h; }
void ref_delay_set_size(void* h, unsigned chans, unsigned length, int fmt) {
  ((SoundDelayBuffer*)h)->SetSize(chans, length, (SampleFormat_t)fmt);
}
unsigned ref_delay_get_channels(void* h) { return ((SoundDelayBuffer*)h)->GetChannels(); }
unsigned ref_delay_get_length(void* h) { return ((SoundDelayBuffer*)h)->GetLength(); }
unsigned ref_delay_get_write_position(void* h) { return ((SoundDelayBuffer*)h)->GetWritePosition(); }
int ref_delay_get_format(void* h) { return (int)((SoundDelayBuffer*)h)->GetFormat(); }
unsigned ref_delay_write_samples(void* h, const void* src, int srcformat, unsigned channel, unsigned nchannels,
                                 unsigned nframes) {
  return ((SoundDelayBuffer*)h)->WriteSamples((const uint8_t*)src, (SampleFormat_t)srcformat, channel, nchannels, nframes);
}
void ref_delay_increment_write_position(void* h, unsigned nframes) {
  ((SoundDelayBuffer*)h)->IncrementWritePosition(nframes);
}
unsigned ref_delay_read_samples(void* h, void* dst, int dstformat, unsigned delay, unsigned channel,
                                unsigned nchannels, unsigned nframes) {
  return ((SoundDelayBuffer*)h)->ReadSamples((uint8_t*)dst, (SampleFormat_t)dstformat, delay, channel, nchannels, nframes);
}
/* raw copy of the ring contents (format-native bytes); returns bytes copied */
unsigned ref_delay_copy_buffer(void* h, void* dst, unsigned maxbytes) {
  SoundDelayBuffer* d = (SoundDelayBuffer*)h;
  unsigned bytes = d->GetChannels() * d->GetLength() * GetBytesPerSample(d->GetFormat());
  const uint8_t* p = NULL;
  const float* pf;
  const double* pd;
  const sint32_t* p32;
  const sint16_t* p16;
  if (d->GetBuffer(&pf)) p = (const uint8_t*)pf;
  else if (d->GetBuffer(&pd)) p = (const uint8_t*)pd;
  else if (d->GetBuffer(&p32)) p = (const uint8_t*)p32;
  else if (d->GetBuffer(&p16)) p = (const uint8_t*)p16;
  if (!p || bytes > maxbytes) return 0;
  memcpy(dst, p, bytes);
  return bytes;
}

/* ---- MultilayerBuffer<float> ("next" row) ---- */
void* ref_mlb_create(unsigned channels, unsigned layers) { return new MultilayerBuffer<float>(channels, layers); }
void ref_mlb_destroy(void* h) { delete (MultilayerBuffer<float>*)h; }
void ref_mlb_write_layer(void* h, unsigned layer, const float* src, unsigned srcchannel, unsigned nsrcchannels,
                         unsigned dstchannel, unsigned nchannels, unsigned nframes) {
  ((MultilayerBuffer<float>*)h)->WriteLayer(layer, src, srcchannel, nsrcchannels, dstchannel, nchannels, nframes, true);
}
unsigned ref_mlb_available_frames(void* h) { return ((MultilayerBuffer<float>*)h)->GetAvailableFrames(); }
unsigned ref_mlb_read_buffer(void* h, unsigned srcchannel, float* dst, unsigned dstchannel, unsigned ndstchannels,
                             unsigned nchannels, unsigned nframes, int overwrite) {
  return ((MultilayerBuffer<float>*)h)->ReadBuffer(srcchannel, dst, dstchannel, ndstchannels, nchannels, nframes, true,
                                                   overwrite != 0);
}


/* ---- BiQuadCoeffs / BiQuad (src/BiQuad.h:27-245, src/BiQuad.cpp:11-497): one coefficient object shared by
 *      nch filters, processed with BiQuad::Process(filters, ...) exactly like BiQuadFilterBank::Process does ---- */
namespace {
struct PeekCoeffs : public BiQuadCoeffs {  // read-only view of the protected interpolation state
  double Mul() const { return mul; }
  double Dec() const { return dec; }
};
struct PeekBiQuad : public BiQuad {
  PeekBiQuad(const BiQuadCoeffs& c) : BiQuad(c) {}
  const double* W() const { return w; }
};
struct RefBank {
  BiQuadCoeffs coeffs;
  std::vector<BiQuad> filters;
  explicit RefBank(unsigned n) : coeffs(), filters(n, BiQuad(coeffs)) {}
};
}  // namespace

void ref_biquad_calc_coeffs(int type, double freq, double fs, double gain, double bandwidth, double* out5) {
  BiQuadCoeffs c((BiQuadCoeffs::Filter_t)type, freq, fs, gain, bandwidth, 0.0);
  out5[0] = c.current.num0;
  out5[1] = c.current.num1;
  out5[2] = c.current.num2;
  out5[3] = c.current.den1;
  out5[4] = c.current.den2;
}
void* ref_biquad_create(unsigned nch) { return new RefBank(nch); }
void ref_biquad_destroy(void* h) { delete (RefBank*)h; }
void ref_biquad_set_coeffs(void* h, const double* c5, double interp_samples) {
  ((RefBank*)h)->coeffs.SetCoeffs(c5[0], c5[1], c5[2], c5[3], c5[4], interp_samples);
}
void ref_biquad_calc(void* h, int type, double freq, double fs, double gain, double bandwidth, double interp_time) {
  ((RefBank*)h)->coeffs.CalcCoeffs((BiQuadCoeffs::Filter_t)type, freq, fs, gain, bandwidth, interp_time);
}
void ref_biquad_process(void* h, const float* src, float* dst, unsigned nchannels, unsigned nsrc, unsigned ndst, unsigned nframes) {
  RefBank* b = (RefBank*)h;
  if (nchannels > b->filters.size()) nchannels = (unsigned)b->filters.size();
  if (!nchannels) return;
  BiQuad::Process(&b->filters[0], src, dst, nchannels, nsrc, ndst, nframes, b->coeffs);
}
void ref_biquad_get_state(void* h, double* w, double* cur5, double* mul_dec) {
  RefBank* b = (RefBank*)h;
  for (size_t j = 0; j < b->filters.size(); j++) {
    const double* ww = static_cast<const PeekBiQuad&>(b->filters[j]).W();
    w[2 * j] = ww[0];
    w[2 * j + 1] = ww[1];
  }
  cur5[0] = b->coeffs.current.num0;
  cur5[1] = b->coeffs.current.num1;
  cur5[2] = b->coeffs.current.num2;
  cur5[3] = b->coeffs.current.den1;
  cur5[4] = b->coeffs.current.den2;
  mul_dec[0] = static_cast<const PeekCoeffs&>(b->coeffs).Mul();
  mul_dec[1] = static_cast<const PeekCoeffs&>(b->coeffs).Dec();
}
void ref_biquad_reset(void* h) {
  RefBank* b = (RefBank*)h;
  for (size_t j = 0; j < b->filters.size(); j++) b->filters[j].Reset();
}

/* ---- BiQuadFilterBank (src/BiQuad.h:247-353, src/BiQuad.cpp:498-662): the reference's own class ---- */
namespace {
struct PeekFbank : public BiQuadFilterBank {
  const BiQuad& Ch(size_t f, size_t j) const { return filters[f]->channels[j]; }
  size_t NCh(size_t f) const { return filters[f]->channels.size(); }
};
}  // namespace
void* ref_fbank_create(unsigned nch, unsigned nfilters) {
  PeekFbank* b = new PeekFbank();
  b->SetChannels(nch);
  b->SetFilters(nfilters);
  return b;
}
void ref_fbank_destroy(void* h) { delete (PeekFbank*)h; }
void ref_fbank_set_filters(void* h, unsigned n) { ((PeekFbank*)h)->SetFilters(n); }
void ref_fbank_add_filter(void* h, const double* c5) {
  BiQuadCoeffs c;
  c.SetCoeffs(c5[0], c5[1], c5[2], c5[3], c5[4], 0.0);
  ((PeekFbank*)h)->AddFilter(c);
}
void ref_fbank_set_channels(void* h, unsigned n) { ((PeekFbank*)h)->SetChannels(n); }
void ref_fbank_set_coeffs(void* h, unsigned filter, const double* c5, double interp_samples) {
  BiQuadCoeffs* c = ((PeekFbank*)h)->GetFilterCoeffs(filter);
  if (c) c->SetCoeffs(c5[0], c5[1], c5[2], c5[3], c5[4], interp_samples);
}
void ref_fbank_calc(void* h, unsigned filter, int type, double freq, double fs, double gain, double bandwidth, double interp_time) {
  BiQuadCoeffs* c = ((PeekFbank*)h)->GetFilterCoeffs(filter);
  if (c) c->CalcCoeffs((BiQuadCoeffs::Filter_t)type, freq, fs, gain, bandwidth, interp_time);
}
void ref_fbank_process(void* h, const float* src, float* dst, unsigned nchannels, unsigned nsrc, unsigned ndst, unsigned nframes) {
  PeekFbank* b = (PeekFbank*)h;
  if (!b->GetChannels()) return;  /* the reference would index an empty vector (src/BiQuad.cpp:653) */
  b->Process(src, dst, nchannels, nsrc, ndst, nframes);
}
void ref_fbank_get_state(void* h, unsigned filter, double* w, double* cur5, double* mul_dec) {
  PeekFbank* b = (PeekFbank*)h;
  if (filter >= b->GetFilters()) return;
  for (size_t j = 0; j < b->NCh(filter); j++) {
    const double* ww = static_cast<const PeekBiQuad&>(b->Ch(filter, j)).W();
    w[2 * j] = ww[0];
    w[2 * j + 1] = ww[1];
  }
  const BiQuadCoeffs* c = b->GetFilterCoeffs(filter);
  cur5[0] = c->current.num0;
  cur5[1] = c->current.num1;
  cur5[2] = c->current.num2;
  cur5[3] = c->current.den1;
  cur5[4] = c->current.den2;
  mul_dec[0] = static_cast<const PeekCoeffs*>(c)->Mul();
  mul_dec[1] = static_cast<const PeekCoeffs*>(c)->Dec();
}
void ref_fbank_reset(void* h) { ((PeekFbank*)h)->Reset(); }


/* ---- AllPassFilterChain<float> (src/AllPassFilter.h:12-262, ring semantics src/RingBuffer.h:17-121) ---- */
namespace {
struct PeekAllPass : public AllPassFilter<float> {
  uint_t Pos() const { return buffer.GetPosition(); }
  const float* Buf() const { return buffer.GetBuffer(); }
  uint_t Len() const { return buffer.GetLength(); }
};
struct PeekChain : public AllPassFilterChain<float> {
  PeekChain(uint_t nch, uint_t nf, const uint_t* d, const float* c) : AllPassFilterChain<float>(nch, nf, d, c) {}
  const PeekAllPass& F(size_t i) const { return static_cast<const PeekAllPass&>(filters[i]); }
  size_t N() const { return filters.size(); }
};
}  // namespace

void* ref_allpass_create(unsigned nchannels, unsigned nfilters, const unsigned* delays, const float* coeffs) {
  return new PeekChain(nchannels, nfilters, delays, coeffs);
}
void ref_allpass_destroy(void* h) { delete (PeekChain*)h; }
void ref_allpass_process(void* h, const float* src, float* dst, unsigned srcchannel, unsigned nsrc, unsigned dstchannel,
                         unsigned ndst, unsigned nframes) {
  ((PeekChain*)h)->Process(src, dst, srcchannel, nsrc, dstchannel, ndst, nframes);
}
/* ring contents of filter f (nchannels * delay floats, raw order) and its write position */
unsigned ref_allpass_get_state(void* h, unsigned f, float* ring, unsigned maxitems) {
  const PeekChain* c = (const PeekChain*)h;
  if (f >= c->N()) return 0;
  const PeekAllPass& a = c->F(f);
  unsigned n = a.Len() < maxitems ? a.Len() : maxitems;
  if (n) memcpy(ring, a.Buf(), n * sizeof(float));
  return a.Pos();
}

/* ---- BiQuadCascade (src/BiQuad.h:373-792): one cascade object per channel of a bank ---- */
namespace {
struct PeekCascade : public BiQuadCascade {
  PeekCascade(uint_t n, bool vec, bool unr) : BiQuadCascade(n, vec, unr) {}
  uint_t N() const { return numfilters; }
  bool Vec() const { return vectorise; }
  void State(float* x12, float* y12, float* w0_12, float* w1_12, float* last) const {
    memcpy(x12, x, sizeof(x));
    memcpy(y12, y, sizeof(y));
    memcpy(w0_12, w0, sizeof(w0));
    memcpy(w1_12, w1, sizeof(w1));
    *last = lastoutput;
  }
};
struct RefCascadeBank {
  std::vector<PeekCascade*> c;
};
}  // namespace

void* ref_cascade_create(unsigned nchannels, unsigned numfilters, int vectorise, int unroll) {
  RefCascadeBank* b = new RefCascadeBank();
  for (unsigned j = 0; j < nchannels; j++) {
    PeekCascade* pc = new PeekCascade(numfilters, vectorise != 0, unroll != 0);
    pc->Reset();  /* the reference's constructor leaves the registers uninitialised */
    b->c.push_back(pc);
  }
  return b;
}
void ref_cascade_destroy(void* h) {
  RefCascadeBank* b = (RefCascadeBank*)h;
  for (size_t j = 0; j < b->c.size(); j++) delete b->c[j];
  delete b;
}
/* channel == ~0u: every channel; returns 1 when the reference accepted the vector */
int ref_cascade_set_coefficients(void* h, unsigned channel, const float* coeffs, unsigned n) {
  RefCascadeBank* b = (RefCascadeBank*)h;
  std::vector<float> v(coeffs, coeffs + n);
  int ok = 1;
  for (size_t j = 0; j < b->c.size(); j++)
    if (channel == ~0u || channel == j) ok &= b->c[j]->SetCoefficients(v) ? 1 : 0;
  return ok;
}
void ref_cascade_reset(void* h) {
  RefCascadeBank* b = (RefCascadeBank*)h;
  for (size_t j = 0; j < b->c.size(); j++) b->c[j]->Reset();
}
/* channel j reads src[j * src_cs + i * src_fs] and writes dst[j * dst_cs + i * dst_fs] through ProcessCascade */
void ref_cascade_process(void* h, const float* src, long src_cs, long src_fs, float* dst, long dst_cs, long dst_fs, unsigned nframes) {
  RefCascadeBank* b = (RefCascadeBank*)h;
  std::vector<float> in(nframes), out(nframes);
  for (size_t j = 0; j < b->c.size(); j++) {
    for (unsigned i = 0; i < nframes; i++) in[i] = src[(long)j * src_cs + (long)i * src_fs];
    if (nframes) b->c[j]->ProcessCascade(&in[0], &out[0], nframes);
    for (unsigned i = 0; i < nframes; i++) dst[(long)j * dst_cs + (long)i * dst_fs] = out[i];
  }
}
/* returns numfilters | vectorise << 8 */
unsigned ref_cascade_get_state(void* h, unsigned channel, float* x12, float* y12, float* w0_12, float* w1_12, float* last) {
  RefCascadeBank* b = (RefCascadeBank*)h;
  if (channel >= b->c.size()) return 0;
  b->c[channel]->State(x12, y12, w0_12, w1_12, last);
  return b->c[channel]->N() | ((unsigned)b->c[channel]->Vec() << 8);
}


}  // extern "C"
