// C++ host-side smoke of the drop-in headers (bbcat-dsp_b200/host/*.h), written the way reference client code
// is written: TransferSamples / MixSamples / Interpolator / FractionalSample / SoundDelayBuffer / Convolver.
// Built and run by tests/test_cpp_host.py on the GPU box; prints PASS and exits 0 on success.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "AllPassFilter.h"
#include "BiQuad.h"
#include "Convolver.h"
#include "FractionalSample.h"
#include "SOFA.h"
#include "SoundDelayBuffer.h"
#include "SoundMixing.h"

using namespace bbcat;

#define CHECK(cond)                                                   \
  do {                                                                \
    if (!(cond)) {                                                    \
      fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                       \
    }                                                                 \
  } while (0)

int main() {
  // TransferSamples: {1..6} 2ch -> channels 1-2 of 4 (reference known answer)
  float src[6] = {1, 2, 3, 4, 5, 6}, dst[12] = {0};
  TransferSamples(src, 0, 2, dst, 1, 4, 2, 3);
  const float want[12] = {0, 1, 2, 0, 0, 3, 4, 0, 0, 5, 6, 0};
  for (int i = 0; i < 12; i++) CHECK(dst[i] == want[i]);
  // float -> 24-bit: +1.0 -> 7fffff, -1e-9 -> ffffff
  float f[2] = {1.0f, -1e-9f};
  uint8_t b[6];
  TransferSamples(f, SampleFormat_Float, false, 0, 2, b, SampleFormat_24bit, false, 0, 2, 2, 1);
  CHECK(b[0] == 0xff && b[1] == 0xff && b[2] == 0x7f && b[3] == 0xff && b[4] == 0xff && b[5] == 0xff);
  // a7: a Ditherer that does nothing (the reference's base class) converts exactly like NULL; a subclass is called once per
  // converted sample with the converter's bit count; TPDFDitherer runs on the device and stays within 1.5 LSB
  {
    struct Count : Ditherer {
      uint_t calls, bits;
      Count() : calls(0), bits(0) {}
      void Dither(uint_t, float& data, uint_t b) { calls++; bits = b; data += 0.0f; }
    };
    float fx[6] = {0.25f, -0.5f, 0.999f, -1.0f, 1e-4f, 0.75f};
    uint8_t plain[18], viabase[18], viacount[18], viatpdf[18];
    Ditherer base;
    Count count;
    TPDFDitherer tpdf;
    TransferSamples(fx, SampleFormat_Float, false, 0, 2, plain, SampleFormat_24bit, false, 0, 2, 2, 3);
    TransferSamples(fx, SampleFormat_Float, false, 0, 2, viabase, SampleFormat_24bit, false, 0, 2, 2, 3, &base);
    TransferSamples(fx, SampleFormat_Float, false, 0, 2, viacount, SampleFormat_24bit, false, 0, 2, 2, 3, &count);
    TransferSamples(fx, SampleFormat_Float, false, 0, 2, viatpdf, SampleFormat_24bit, false, 0, 2, 2, 3, &tpdf);
    CHECK(memcmp(plain, viabase, 18) == 0 && memcmp(plain, viacount, 18) == 0);
    CHECK(count.calls == 6 && count.bits == 8);
    for (int i = 0; i < 6; i++) {
      const int a = (int)((uint32_t)plain[3 * i] | ((uint32_t)plain[3 * i + 1] << 8) | ((uint32_t)(int8_t)plain[3 * i + 2] << 16));
      const int b = (int)((uint32_t)viatpdf[3 * i] | ((uint32_t)viatpdf[3 * i + 1] << 8) | ((uint32_t)(int8_t)viatpdf[3 * i + 2] << 16));
      CHECK(abs(a - b) <= 2);
    }
    // widening conversions never call the ditherer
    float wide[2];
    sint16_t narrow[2] = {1000, -1000};
    count.calls = 0;
    TransferSamples(narrow, 0, 2, wide, 0, 2, 2, 1, &count);
    CHECK(count.calls == 0 && wide[0] == 1000.0f / 32768.0f);
  }
  // MixSamples with an Interpolator: gains 0, .25, .5 and the object ends at .75
  float ones[3] = {1, 1, 1}, acc[3] = {0, 0, 0};
  Interpolator interp(1.0f, 0.0f);
  MixSamples(ones, 0, 1, acc, 0, 1, 1, 3, interp, 0.25f);
  CHECK(acc[0] == 0.0f && acc[1] == 0.25f && acc[2] == 0.5f && (float)interp == 0.75f);
  // FractionalSample KAT
  std::vector<float> ring(64, 0.0f);
  ring[20] = 1.0f;
  CHECK(fabs(FractionalSample(&ring[0], 0, 1, 64, 28.0) - 9.976246356964e-01) < 1e-12);
  CHECK(FractionalSampleAdditionalDelayRequired() == 14);
  // SoundDelayBuffer
  SoundDelayBuffer delay;
  delay.SetSize(2, 8, SampleFormat_Float);
  float frames[12];
  for (int i = 0; i < 12; i++) frames[i] = (float)i;
  CHECK(delay.WriteSamples(frames, 0, 2, 6) == 6);
  delay.IncrementWritePosition(6);
  float rd[8];
  CHECK(delay.ReadSamples(rd, 4, 0, 2, 4) == 4);
  CHECK(rd[0] == 4.0f && rd[7] == 11.0f);
  // Convolver: 2 channels, pure-delay IRs, 3 blocks of 64
  const uint_t B = 64;
  Convolver conv(B, 2, 2, 3);
  std::vector<float> ir0(100, 0.0f), ir1(100, 0.0f);
  ir0[0] = 1.0f;
  ir1[70] = 0.5f;
  ConvolverFilter* f0 = conv.CreateFilter(&ir0[0], 100);
  ConvolverFilter* f1 = conv.CreateFilter(&ir1[0], 100);
  conv.SelectFilter(0, f0);
  conv.SelectFilter(1, f1);
  std::vector<float> x(3 * B * 2), y(3 * B * 2, 0.0f);
  for (size_t i = 0; i < x.size(); i++) x[i] = (float)((i * 7919u) % 1000u) / 1000.0f - 0.5f;
  conv.Convolve(&x[0], 2, &y[0], 2, 3 * B);
  double err = 0;
  for (uint_t n = 0; n < 3 * B; n++) {
    err = fmax(err, fabs(y[2 * n] - x[2 * n]));
    double w1 = n >= 70 ? 0.5 * x[2 * (n - 70) + 1] : 0.0;
    err = fmax(err, fabs(y[2 * n + 1] - w1));
  }
  CHECK(err < 5e-6);
  delete f0;
  delete f1;
  // BiQuadBank: y[n] = x[n] + 0.5 x[n-1] - 0.25 y[n-1] on an impulse (exact in binary)
  {
    BiQuadBank bank(2);
    bank.SetCoeffs(1.0, 0.5, 0.0, 0.25, 0.0);
    float bx[8] = {1, 0, 0, 0, 0, 0, 0, 0}, by[8] = {9, 9, 9, 9, 9, 9, 9, 9};
    bank.Process(bx, by, 1, 2, 2, 4);
    CHECK(by[0] == 1.0f && by[2] == 0.25f && by[4] == -0.0625f && by[6] == 0.015625f && by[1] == 9.0f);
    bank.CalcCoeffs(BiQuadBank::FLAT, 1000.0, 48000.0);
    CHECK(bank.GetCurrent().num0 == 1.0 && bank.GetCurrent().den1 == 0.0);
  }
  // BiQuadFilterBank: two of those filters in series equal two passes of one (exact), and AddFilter / GetFilterCoeffs work
  {
    BiQuadFilterBank fb;
    fb.SetChannels(2);
    fb.SetFilters(1);
    fb.GetFilterCoeffs(0).SetCoeffs(1.0, 0.5, 0.0, 0.25, 0.0);
    BiQuadFilterBank::COEFFS c2 = {1.0, 0.5, 0.0, 0.25, 0.0};
    fb.AddFilter(c2);
    CHECK(fb.GetFilters() == 2 && fb.GetChannels() == 2 && !fb.GetFilterCoeffs(2).Valid());
    BiQuadBank one(2), two(2);
    one.SetCoeffs(1.0, 0.5, 0.0, 0.25, 0.0);
    two.SetCoeffs(1.0, 0.5, 0.0, 0.25, 0.0);
    float bx[16], by[16], bz[16];
    for (int i = 0; i < 16; i++) bx[i] = (float)((i * 7) % 5) - 2.0f, by[i] = 9.0f;
    fb.Process(bx, by, 2, 2, 2, 8);
    one.Process(bx, bz, 2, 2, 2, 8);
    two.Process(bz, bz, 2, 2, 2, 8);
    CHECK(memcmp(by, bz, sizeof(by)) == 0);
    CHECK(fb.GetFilterCoeffs(1).GetCurrent().num1 == 0.5);
  }
  // SoundRingBuffer: one frame always stays free; writes are limited by the read position
  {
    SoundRingBuffer ring;
    ring.SetSize(1, 8, SampleFormat_Float);
    float rx[10] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
    CHECK(ring.GetWriteFramesAvailable() == 7 && ring.WriteSamples(rx, 0, 1, 10) == 7);
    ring.IncrementWritePosition(10);
    CHECK(ring.GetWritePosition() == 7 && ring.GetReadFramesAvailable() == 7 && ring.GetWriteFramesAvailable() == 0);
    ring.IncrementReadPosition(3);
    float ry[8] = {0};
    CHECK(ring.ReadSamples(ry, 2, 0, 1, 8) == 2 && ry[0] == 6.0f && ry[1] == 7.0f);
  }
  // BiQuadCascadeBank: one section per channel, b1 = 1 -> y[n] = x[n] + x[n-1]; a wrong-length vector is refused
  {
    BiQuadCascadeBank casc(2, 1, false);
    const float cf[5] = {7.0f, 1.0f, 0.0f, 0.0f, 0.0f};
    CHECK(casc.SetCoefficients(cf, 5));
    CHECK(!casc.SetCoefficients(cf, 4));
    float cx[8] = {1, 10, 2, 20, 3, 30, 4, 40}, cy[8] = {0};
    casc.ProcessCascade(cx, cy, 4);
    CHECK(cy[0] == 1.0f && cy[1] == 10.0f && cy[2] == 3.0f && cy[3] == 30.0f && cy[6] == 7.0f && cy[7] == 70.0f);
  }
  // AllPassFilterChain: one section, delay 2, c = 0.5, impulse -> 0.5, 0, 0.75, 0, -0.375
  {
    const uint_t d[1] = {2};
    const float c[1] = {0.5f};
    AllPassFilterChain chain(1, 1, d, c);
    float ax[6] = {1, 0, 0, 0, 0, 0}, ay[6] = {0};
    chain.Process(ax, ay, 0, 1, 0, 1, 6);
    CHECK(ay[0] == 0.5f && ay[1] == 0.0f && ay[2] == 0.75f && ay[3] == 0.0f && ay[4] == -0.375f);
  }
  // SOFA: a file that does not exist throws with the library's message (reading real sets: tests/test_sofa.py)
  {
    bool threw = false;
    try {
      SOFA sofa("/nonexistent/set.sofa");
    } catch (const std::runtime_error& e) {
      threw = strstr(e.what(), "cannot open") != NULL;
    }
    CHECK(threw);
  }
  printf("PASS max_err=%g\n", err);
  return 0;
}
