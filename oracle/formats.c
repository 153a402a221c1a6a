/* TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of the reference's sample-format layer:
 *   GetBitsPerSample / GetBytesPerSample   SoundFormatConversions.cpp:14-40
 *   BlockTransferSanityChecks              SoundFormatConversions.cpp:59-93
 *   TransferSamples                        SoundFormatConversions.cpp:151-198
 *   TransferSamplesLinear                  SoundFormatConversions.cpp:204-219
 *   the 2x2x6x6 converter table            SoundFormatRawConversions.cpp:4516-4869
 *
 * The 96 generated converters all follow one law (genconversions.php), restated
 * here as "load one sample into its class register, convert class, store":
 *   integer source  : bytes -> sint32 with the value in the TOP bits          (.cpp:98-103, 393, 587)
 *   int -> int      : keep the top bytes                                      (.cpp:104-107, 303)
 *   int -> float    : (float)sval * 2^-31f   /  int -> double: (double)sval * 2^-31   (.cpp:166,199 / :220,253)
 *   float|double -> int : (sint32) clamp(x * 2^31 [double], -2147483648.0, 2147483647.0), C truncation
 *                                                                              (.cpp:701,751,795 / :904,954,1000)
 *   float <-> double: C casts                                                  (.cpp:849, 1050)
 *   same format     : memcpy per frame when endianness matches (.cpp:20-62), byte swap otherwise (.cpp:1068+)
 * Iteration order is restated too (it only matters for dst == src): frames run
 * backwards when the destination frame is longer (SoundFormatConversions.cpp:178-185)
 * and channels run backwards inside a frame when the destination sample is wider
 * (e.g. SoundFormatRawConversions.cpp:180-190).
 */
#include "oracle.h"

#include <string.h>

static const unsigned char k_bits[ORC_FMT_COUNT] = {1, 16, 24, 32, 32, 64};

unsigned orc_get_bits_per_sample(int fmt) { return k_bits[fmt]; }
unsigned orc_get_bytes_per_sample(int fmt) { return (k_bits[fmt] + 7u) >> 3; }

static unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }

int orc_block_transfer_sanity_checks(unsigned* src_channel, unsigned* src_channels, unsigned* dst_channel,
                                     unsigned* dst_channels, unsigned* nchannels, unsigned* nframes,
                                     int allowsinglechannel) {
  if (!(*src_channels && *dst_channels && *nframes && *nchannels)) return 0;
  *src_channel = umin(*src_channel, *src_channels - 1);
  *dst_channel = umin(*dst_channel, *dst_channels - 1);
  *nchannels = umin(*nchannels, *src_channels - *src_channel);
  *nchannels = umin(*nchannels, *dst_channels - *dst_channel);
  if (!*nchannels) return 0;
  if (allowsinglechannel && *nchannels == *src_channels && *nchannels == *dst_channels) {
    /* contiguous both sides: one frame of many channels (.cpp:81-86) */
    *nchannels *= *nframes;
    *nframes = 1;
  }
  return 1;
}

/* one sample held in the register class of its format */
typedef struct {
  int32_t i; /* integer formats: value in the top bits */
  float f;
  double d;
} sample_reg;

static void load_sample(const uint8_t* p, int fmt, int be, sample_reg* r) {
  uint8_t b[8];
  unsigned n = orc_get_bytes_per_sample(fmt), k;
  /* b[] = little-endian byte order of the sample */
  for (k = 0; k < n; k++) b[k] = be ? p[n - 1 - k] : p[k];
  switch (fmt) {
    case ORC_FMT_16BIT: r->i = (int32_t)(((uint32_t)b[1] << 24) + ((uint32_t)b[0] << 16)); break;
    case ORC_FMT_24BIT: r->i = (int32_t)(((uint32_t)b[2] << 24) + ((uint32_t)b[1] << 16) + ((uint32_t)b[0] << 8)); break;
    case ORC_FMT_32BIT:
      r->i = (int32_t)(((uint32_t)b[3] << 24) + ((uint32_t)b[2] << 16) + ((uint32_t)b[1] << 8) + (uint32_t)b[0]);
      break;
    case ORC_FMT_FLOAT: memcpy(&r->f, b, 4); break;
    default: memcpy(&r->d, b, 8); break;
  }
}

static void store_bytes(uint8_t* p, const uint8_t* le, unsigned n, int be) {
  unsigned k;
  for (k = 0; k < n; k++) p[k] = be ? le[n - 1 - k] : le[k];
}

static int32_t float_to_int(double scaled) {
  /* limited::limit(v, -2147483648.0, 2147483647.0) then C truncation (.cpp:701) */
  double lo = -2147483648.0, hi = 2147483647.0;
  double v = (scaled < lo) ? lo : scaled; /* std::max(v, lo): returns v when !(v < lo), incl. NaN */
  v = (hi < v) ? hi : v;                  /* std::min(v, hi) */
  if (v != v) return INT32_MIN;           /* NaN -> x86 cvttsd2si "integer indefinite"; parity-unpinned (UB in C) */
  return (int32_t)v;
}

static void convert_store(const sample_reg* r, int srcfmt, uint8_t* p, int dstfmt, int be) {
  uint8_t le[8];
  int src_is_int = (srcfmt <= ORC_FMT_32BIT);
  if (dstfmt <= ORC_FMT_32BIT) {
    int32_t v;
    if (src_is_int) v = r->i;
    else if (srcfmt == ORC_FMT_FLOAT) v = float_to_int((double)r->f * 2147483648.0);
    else v = float_to_int(r->d * 2147483648.0);
    le[0] = (uint8_t)((uint32_t)v);
    le[1] = (uint8_t)((uint32_t)v >> 8);
    le[2] = (uint8_t)((uint32_t)v >> 16);
    le[3] = (uint8_t)((uint32_t)v >> 24);
    if (dstfmt == ORC_FMT_16BIT) store_bytes(p, le + 2, 2, be);
    else if (dstfmt == ORC_FMT_24BIT) store_bytes(p, le + 1, 3, be);
    else store_bytes(p, le, 4, be);
  } else if (dstfmt == ORC_FMT_FLOAT) {
    float v;
    if (src_is_int) v = (float)r->i * 4.656612873077392578125e-10f; /* 2^-31 */
    else if (srcfmt == ORC_FMT_FLOAT) v = r->f;
    else v = (float)r->d;
    memcpy(le, &v, 4);
    store_bytes(p, le, 4, be);
  } else {
    double v;
    if (src_is_int) v = (double)r->i * 4.656612873077392578125e-10; /* 2^-31 */
    else if (srcfmt == ORC_FMT_FLOAT) v = (double)r->f;
    else v = r->d;
    memcpy(le, &v, 8);
    store_bytes(p, le, 8, be);
  }
}

/* a7: the converters that call ditherer->Dither(i, sval, bits) and the bit count they pass (SoundFormatRawConversions.cpp:
 * 24/32 -> 16 and float/double -> 16: 16 (:301, :499, :699, :902); 32/float/double -> 24: 8 (:541, :749, :952);
 * double -> 32: 0 (:998)); -1 = no call site */
int orc_dither_bits(int src, int dst) {
  if (dst == ORC_FMT_16BIT && (src == ORC_FMT_24BIT || src == ORC_FMT_32BIT || src == ORC_FMT_FLOAT || src == ORC_FMT_DOUBLE)) return 16;
  if (dst == ORC_FMT_24BIT && (src == ORC_FMT_32BIT || src == ORC_FMT_FLOAT || src == ORC_FMT_DOUBLE)) return 8;
  if (dst == ORC_FMT_32BIT && src == ORC_FMT_DOUBLE) return 0;
  return -1;
}

/* Dither_TPDF as libbbx defines it (include/bbx.h, a7): there is no reference implementation, this is the second
 * statement of OUR law, so GPU-vs-oracle equality checks the kernel, not BBC parity. */
static uint64_t dither_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void dither_tpdf(sample_reg* r, int srcfmt, int bits, uint64_t seed, uint64_t index) {
  const uint64_t z = dither_mix64(seed ^ (index * 0x9E3779B97F4A7C15ull));
  const uint32_t u1 = (uint32_t)z, u2 = (uint32_t)(z >> 32);
  if (srcfmt <= ORC_FMT_32BIT) {
    int64_t v = (int64_t)r->i + (int64_t)(u1 >> (32 - bits)) - (int64_t)(u2 >> (32 - bits)) + ((int64_t)1 << (bits - 1));
    if (v > 2147483647ll) v = 2147483647ll;
    if (v < -2147483648ll) v = -2147483648ll;
    r->i = (int32_t)v;
  } else {
    const double t = ((double)u1 - (double)u2) * 2.3283064365386962890625e-10 + (bits > 0 ? 0.5 : 0.0);
    const double n = t * ((double)(1u << bits) * 4.656612873077392578125e-10);
    if (srcfmt == ORC_FMT_FLOAT) r->f = (float)((double)r->f + n);
    else r->d = r->d + n;
  }
}

void orc_transfer_samples(const void* vsrc, int srctype, int src_be, unsigned src_channel, unsigned src_channels,
                          void* vdst, int dsttype, int dst_be, unsigned dst_channel, unsigned dst_channels,
                          unsigned nchannels, unsigned nframes) {
  orc_transfer_samples_dither(vsrc, srctype, src_be, src_channel, src_channels, vdst, dsttype, dst_be, dst_channel, dst_channels,
                              nchannels, nframes, 0, 0, 0, 0);
}

/* dither: 0 none, 1 TPDF(seed).  hook != NULL restates the reference's Ditherer call: hook(user, i, kind, &sval, bits) with
 * i = the frame LOOP counter (it counts up even when the frames run backwards), kind 0 sint32 / 1 float / 2 double */
void orc_transfer_samples_dither(const void* vsrc, int srctype, int src_be, unsigned src_channel, unsigned src_channels,
                                 void* vdst, int dsttype, int dst_be, unsigned dst_channel, unsigned dst_channels,
                                 unsigned nchannels, unsigned nframes, int dither, uint64_t seed, orc_dither_hook hook,
                                 void* user) {
  if (!orc_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 1))
    return;
  if (srctype <= ORC_FMT_UNKNOWN || srctype >= ORC_FMT_COUNT || dsttype <= ORC_FMT_UNKNOWN || dsttype >= ORC_FMT_COUNT)
    return;
  src_be = src_be != 0;
  dst_be = dst_be != 0;

  const uint8_t* src = (const uint8_t*)vsrc;
  uint8_t* dst = (uint8_t*)vdst;
  int srclen = (int)orc_get_bytes_per_sample(srctype), dstlen = (int)orc_get_bytes_per_sample(dsttype);
  long srcflen = (long)src_channels * srclen, dstflen = (long)dst_channels * dstlen;
  unsigned i, j;

  src += (size_t)src_channel * srclen;
  dst += (size_t)dst_channel * dstlen;
  int reversed = 0;
  if (dstflen > srcflen) { /* bigger destination rectangle: run from the last frame backwards */
    src += (long)(nframes - 1) * srcflen;
    dst += (long)(nframes - 1) * dstlen * dst_channels;
    srcflen = -srcflen;
    dstflen = -dstflen;
    reversed = 1;
  }
  const int dbits = orc_dither_bits(srctype, dsttype);

  if (srctype == dsttype && src_be == dst_be) { /* __CopyMemory_n */
    for (i = 0; i < nframes; i++, src += srcflen, dst += dstflen)
      if (dst != src) memcpy(dst, src, (size_t)nchannels * srclen);
    return;
  }

  /* channel direction inside a frame: backwards when the destination format index is higher
   * (wider or equal-width-later sample), genconversions.php:156-167 */
  int backwards = (srctype < dsttype);
  for (i = 0; i < nframes; i++, src += srcflen, dst += dstflen) {
    for (j = 0; j < nchannels; j++) {
      unsigned c = backwards ? (nchannels - 1 - j) : j;
      sample_reg r;
      load_sample(src + (size_t)c * srclen, srctype, src_be, &r);
      if (dbits >= 0) {
        if (hook) {
          if (srctype <= ORC_FMT_32BIT) hook(user, i, 0, &r.i, (unsigned)dbits);
          else if (srctype == ORC_FMT_FLOAT) hook(user, i, 1, &r.f, (unsigned)dbits);
          else hook(user, i, 2, &r.d, (unsigned)dbits);
        } else if (dither == 1) {
          /* our law indexes the sample by its position in the rectangle, whatever the loop direction */
          const unsigned frame = reversed ? nframes - 1 - i : i;
          dither_tpdf(&r, srctype, dbits, seed, (uint64_t)frame * nchannels + c);
        }
      }
      convert_store(&r, srctype, dst + (size_t)c * dstlen, dsttype, dst_be);
    }
  }
}

void orc_transfer_samples_linear(const void* vsrc, int srctype, void* vdst, int dsttype, unsigned nsamples) {
  /* one frame of nsamples channels, machine endianness both sides (.cpp:204-219) */
  if (srctype < 0 || srctype >= ORC_FMT_COUNT || dsttype < 0 || dsttype >= ORC_FMT_COUNT) return;
  if (srctype == ORC_FMT_UNKNOWN || dsttype == ORC_FMT_UNKNOWN) return; /* NULL table entry -> BBCERROR + return */
  unsigned j;
  int srclen = (int)orc_get_bytes_per_sample(srctype), dstlen = (int)orc_get_bytes_per_sample(dsttype);
  const uint8_t* src = (const uint8_t*)vsrc;
  uint8_t* dst = (uint8_t*)vdst;
  if (srctype == dsttype) {
    if (dst != src) memcpy(dst, src, (size_t)nsamples * srclen);
    return;
  }
  int backwards = (srctype < dsttype);
  for (j = 0; j < nsamples; j++) {
    unsigned c = backwards ? (nsamples - 1 - j) : j;
    sample_reg r;
    load_sample(src + (size_t)c * srclen, srctype, 0, &r);
    convert_store(&r, srctype, dst + (size_t)c * dstlen, dsttype, 0);
  }
}

/* tests/cpp/test_ditherer.h restated in C: the stateful test Ditherer the reference wrapper and the shim test run */
static uint32_t test_pattern(uint32_t* calls, unsigned channel) { return ((*calls)++ * 40503u + channel * 7919u) * 2654435761u >> 8; }
void orc_test_dither_hook(void* user, unsigned channel, int kind, void* data, unsigned bits) {
  uint32_t* calls = (uint32_t*)user;
  if (kind == 0) {
    const uint32_t mask = bits ? ((1u << bits) - 1u) : 0u;
    int64_t v = (int64_t)(*(int32_t*)data) + (int64_t)(test_pattern(calls, channel) & mask) - (int64_t)(mask >> 1);
    if (v > 2147483647ll) v = 2147483647ll;
    if (v < -2147483648ll) v = -2147483648ll;
    *(int32_t*)data = (int32_t)v;
  } else {
    const double n = ((double)(test_pattern(calls, channel) & 0xffffu) / 32768.0 - 1.0) * (double)(1u << bits) / 2147483648.0;
    if (kind == 1) *(float*)data += (float)n;
    else *(double*)data += n;
  }
}

/* same signature as ref_transfer_samples_ditherer (oracle/ref_wrap.cpp): mode 0 = a no-op Ditherer, 1 = the test Ditherer */
static void noop_hook(void* user, unsigned channel, int kind, void* data, unsigned bits) {
  (void)user; (void)channel; (void)kind; (void)data; (void)bits;
}
unsigned orc_transfer_samples_ditherer(const void* src, int srctype, int src_be, unsigned src_channel, unsigned src_channels,
                                       void* dst, int dsttype, int dst_be, unsigned dst_channel, unsigned dst_channels,
                                       unsigned nchannels, unsigned nframes, int mode) {
  uint32_t calls = 0;
  orc_transfer_samples_dither(src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel, dst_channels,
                              nchannels, nframes, 0, 0, mode ? orc_test_dither_hook : noop_hook, &calls);
  return calls;
}
