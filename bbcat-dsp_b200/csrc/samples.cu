// samples.cu -- sample-format, mixing, fractional-read and delay-ring entry points of libbbx.
//
// GPU counterparts of the in-tree bbcat-dsp pieces that sit either side of the convolver:
//   TransferSamples / TransferSamplesLinear   src/SoundFormatConversions.cpp:151-219
//   MixSamples (plain and interpolated)       src/SoundMixing.h:55-81, src/SoundMixing.cpp:23-52
//   FractionalSample                          src/FractionalSample.cpp:249-341
//   SoundDelayBuffer                          src/SoundDelayBuffer.cpp:11-191
// Integer/byte results are bit-exact with the reference; float results too, because every
// product and sum is rounded separately (__fmul_rn/__fadd_rn, no FMA contraction).
#include <stdarg.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "formats.cuh"
#include "fracsample.cuh"

namespace bbx {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no usable CUDA device (%s); libbbx has no CPU fallback", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    return BBX_ERR_CUDA;
  }
  return BBX_OK;
}

int DeviceScratch::ensure(size_t bytes) {
  int dev = 0;
  BBX_CUDA_TRY(cudaGetDevice(&dev));
  if (ptr && (dev != device || cap < bytes)) {
    cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  if (!ptr) {
    size_t want = std::max<size_t>(bytes, 1 << 16);
    BBX_CUDA_TRY(cudaMalloc(&ptr, want));
    cap = want;
    device = dev;
  }
  return BBX_OK;
}
DeviceScratch::~DeviceScratch() {
  // process teardown: the context may already be gone, ignore errors
  if (ptr) cudaFree(ptr);
}
DeviceScratch& scratch(int which) {
  // one set per device: objects on different GPUs used from one thread do not evict each other's scratch
  static thread_local DeviceScratch s[16][4];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return s[dev & 15][which & 3];
}

// ------------------------------------------------------------------------------------------
// TransferSamples kernel: one thread per (frame, channel) of the rectangle
// ------------------------------------------------------------------------------------------
struct XferGeom {
  const uint8_t* src;
  uint8_t* dst;
  uint32_t nchannels, nframes;
  uint64_t src_frame_bytes, dst_frame_bytes;  // frame strides
  bool src_be, dst_be;
  int dither;          // bbx_dither: 0 none, 1 TPDF (only the converters with a dither call site look at it)
  uint64_t seed;
};

template <int SRC, int DST, bool ALIGNED>
__global__ void __launch_bounds__(256) k_transfer(XferGeom g) {
  const uint64_t total = (uint64_t)g.nchannels * g.nframes;
  const uint32_t srclen = fmt_bytes(SRC), dstlen = fmt_bytes(DST);
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t frame = (uint32_t)(idx / g.nchannels), ch = (uint32_t)(idx % g.nchannels);
    const uint8_t* sp = g.src + frame * g.src_frame_bytes + (uint64_t)ch * srclen;
    uint8_t* dp = g.dst + frame * g.dst_frame_bytes + (uint64_t)ch * dstlen;
    SampleReg r = ALIGNED ? load_sample_aligned_le<SRC>(sp) : load_sample<SRC>(sp, g.src_be);
    uint64_t bits;
    if (SRC == DST) {
      // same format: raw copy / byte swap, no value conversion (NaN payloads survive), .cpp:20-62
      bits = (SRC <= FMT_32) ? (uint64_t)((uint32_t)r.i >> (32 - 8 * fmt_bytes(SRC)))
             : (SRC == FMT_F32) ? (uint64_t)__float_as_uint(r.f) : (uint64_t)__double_as_longlong(r.d);
    } else {
      if constexpr (dither_bits(SRC, DST) >= 0) {
        if (g.dither == BBX_DITHER_TPDF) dither_tpdf<SRC, dither_bits(SRC, DST)>(r, g.seed, idx);
      }
      bits = convert_sample<SRC, DST>(r);
    }
    if (ALIGNED) store_sample_aligned_le<DST>(dp, bits);
    else store_sample<DST>(dp, bits, g.dst_be);
  }
}

template <int SRC, int DST>
static void launch_transfer2(const XferGeom& g, bool aligned, cudaStream_t st) {
  uint64_t total = (uint64_t)g.nchannels * g.nframes;
  uint32_t blocks = (uint32_t)std::min<uint64_t>((total + 255) / 256, 148u * 16u);
  if (aligned) k_transfer<SRC, DST, true><<<blocks, 256, 0, st>>>(g);
  else k_transfer<SRC, DST, false><<<blocks, 256, 0, st>>>(g);
}

template <int SRC>
static void launch_transfer1(int dst, const XferGeom& g, bool aligned, cudaStream_t st) {
  switch (dst) {
    case FMT_16: launch_transfer2<SRC, FMT_16>(g, aligned, st); break;
    case FMT_24: launch_transfer2<SRC, FMT_24>(g, aligned, st); break;
    case FMT_32: launch_transfer2<SRC, FMT_32>(g, aligned, st); break;
    case FMT_F32: launch_transfer2<SRC, FMT_F32>(g, aligned, st); break;
    default: launch_transfer2<SRC, FMT_F64>(g, aligned, st); break;
  }
}

static bool is_aligned(const void* p, uint64_t stride, uint32_t len) {
  if (len == 3) return true;  // byte path either way
  return ((uintptr_t)p % len) == 0 && (stride % len) == 0;
}

// Device-pointer rectangle transfer after the sanity checks have been applied.
static int transfer_dev_checked(const void* src, int srctype, bool src_be, uint32_t src_channel, uint32_t src_channels,
                                void* dst, int dsttype, bool dst_be, uint32_t dst_channel, uint32_t dst_channels,
                                uint32_t nchannels, uint32_t nframes, cudaStream_t st, int dither = 0, uint64_t seed = 0) {
  XferGeom g;
  g.dither = dither;
  g.seed = seed;
  uint32_t srclen = fmt_bytes(srctype), dstlen = fmt_bytes(dsttype);
  g.src = (const uint8_t*)src + (uint64_t)src_channel * srclen;
  g.dst = (uint8_t*)dst + (uint64_t)dst_channel * dstlen;
  g.nchannels = nchannels;
  g.nframes = nframes;
  g.src_frame_bytes = (uint64_t)src_channels * srclen;
  g.dst_frame_bytes = (uint64_t)dst_channels * dstlen;
  g.src_be = src_be;
  g.dst_be = dst_be;
  bool aligned = !src_be && !dst_be && srctype != FMT_24 && dsttype != FMT_24 &&
                 is_aligned(g.src, g.src_frame_bytes, srclen) && is_aligned(g.dst, g.dst_frame_bytes, dstlen);
  switch (srctype) {
    case FMT_16: launch_transfer1<FMT_16>(dsttype, g, aligned, st); break;
    case FMT_24: launch_transfer1<FMT_24>(dsttype, g, aligned, st); break;
    case FMT_32: launch_transfer1<FMT_32>(dsttype, g, aligned, st); break;
    case FMT_F32: launch_transfer1<FMT_F32>(dsttype, g, aligned, st); break;
    default: launch_transfer1<FMT_F64>(dsttype, g, aligned, st); break;
  }
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

static bool valid_fmt(int f) { return f > FMT_UNKNOWN && f < FMT_COUNT; }

// host rectangle <-> compact device buffer [nframes][nchannels]
static int copy_rect_h2d(void* dcompact, const uint8_t* hsrc, uint32_t channel, uint32_t channels, uint32_t nchannels,
                         uint32_t nframes, uint32_t len, cudaStream_t st) {
  size_t width = (size_t)nchannels * len, pitch = (size_t)channels * len;
  const uint8_t* p = hsrc + (size_t)channel * len;
  if (width == pitch || nframes == 1) BBX_CUDA_TRY(cudaMemcpyAsync(dcompact, p, width * nframes, cudaMemcpyHostToDevice, st));
  else BBX_CUDA_TRY(cudaMemcpy2DAsync(dcompact, width, p, pitch, width, nframes, cudaMemcpyHostToDevice, st));
  return BBX_OK;
}
static int copy_rect_d2h(uint8_t* hdst, const void* dcompact, uint32_t channel, uint32_t channels, uint32_t nchannels,
                         uint32_t nframes, uint32_t len, cudaStream_t st) {
  size_t width = (size_t)nchannels * len, pitch = (size_t)channels * len;
  uint8_t* p = hdst + (size_t)channel * len;
  if (width == pitch || nframes == 1) BBX_CUDA_TRY(cudaMemcpyAsync(p, dcompact, width * nframes, cudaMemcpyDeviceToHost, st));
  else BBX_CUDA_TRY(cudaMemcpy2DAsync(p, pitch, dcompact, width, width, nframes, cudaMemcpyDeviceToHost, st));
  return BBX_OK;
}

// ------------------------------------------------------------------------------------------
// MixSamples kernels
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

// dst[f][j] += gain(f) * src[f][j];  gains == nullptr -> constant mul
template <typename T>
__global__ void __launch_bounds__(256) k_mix(const T* __restrict__ src, uint32_t src_channels, T* __restrict__ dst,
                                             uint32_t dst_channels, uint32_t nchannels, uint32_t nframes, T mul,
                                             const float* __restrict__ gains) {
  const uint64_t total = (uint64_t)nchannels * nframes;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t frame = (uint32_t)(idx / nchannels), ch = (uint32_t)(idx % nchannels);
    T g = gains ? (T)gains[frame] : mul;
    T s = src[(uint64_t)frame * src_channels + ch];
    T* d = dst + (uint64_t)frame * dst_channels + ch;
    *d = add_rn<T>(*d, mul_rn<T>(g, s));
  }
}

// Interpolator::operator+= (src/Interpolator.h:55)
static inline float interp_step(float target, float current, float inc) {
  return (target >= current) ? std::min(current + inc, target) : std::max(current - inc, target);
}

// ------------------------------------------------------------------------------------------
// FractionalSample
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_frac(const T* __restrict__ buffer, uint32_t channel, uint32_t channels,
                                              uint32_t length, const double* __restrict__ pos, uint32_t n,
                                              double* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fractional_sample_dev<T>(buffer, channel, channels, length, pos[i]);
}

}  // namespace bbx

using namespace bbx;

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int bbx_version(void) { return BBX_VERSION; }
const char* bbx_last_error(void) { return get_error(); }

int bbx_device_count(int* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (count) *count = (e == cudaSuccess) ? n : 0;
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    return BBX_ERR_CUDA;
  }
  return BBX_OK;
}

int bbx_host_alloc(void** ptr, size_t bytes) {
  BBX_REQUIRE(ptr != nullptr, "bbx_host_alloc: null out pointer");
  BBX_CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return BBX_OK;
}
int bbx_host_free(void* ptr) {
  if (ptr) BBX_CUDA_TRY(cudaFreeHost(ptr));
  return BBX_OK;
}

int bbx_shard_range(uint32_t nchannels, uint32_t rank, uint32_t world, uint32_t* first, uint32_t* count) {
  BBX_REQUIRE(world > 0 && rank < world && first && count, "bbx_shard_range: bad rank/world");
  // block-contiguous: the first (nchannels % world) ranks take one extra channel
  uint32_t base = nchannels / world, extra = nchannels % world;
  *count = base + (rank < extra ? 1u : 0u);
  *first = rank * base + std::min(rank, extra);
  return BBX_OK;
}

uint8_t bbx_get_bits_per_sample(int format) {
  static const uint8_t bits[FMT_COUNT] = {1, 16, 24, 32, 32, 64};
  return (format >= 0 && format < FMT_COUNT) ? bits[format] : 0;
}
uint8_t bbx_get_bytes_per_sample(int format) { return (uint8_t)((bbx_get_bits_per_sample(format) + 7) >> 3); }

int bbx_block_transfer_sanity_checks(uint32_t* src_channel, uint32_t* src_channels, uint32_t* dst_channel,
                                     uint32_t* dst_channels, uint32_t* nchannels, uint32_t* nframes,
                                     int allowsinglechannel) {
  // src/SoundFormatConversions.cpp:59-93
  if (!*src_channels || !*dst_channels || !*nframes || !*nchannels) return 0;
  *src_channel = std::min(*src_channel, *src_channels - 1);
  *dst_channel = std::min(*dst_channel, *dst_channels - 1);
  *nchannels = std::min(*nchannels, *src_channels - *src_channel);
  *nchannels = std::min(*nchannels, *dst_channels - *dst_channel);
  if (!*nchannels) return 0;
  if (allowsinglechannel && *nchannels == *src_channels && *nchannels == *dst_channels) {
    *nchannels *= *nframes;  // both sides contiguous: a single frame of many channels
    *nframes = 1;
  }
  return 1;
}

int bbx_dither_bits(int srctype, int dsttype) { return dither_bits(srctype, dsttype); }

int bbx_transfer_samples_dev(const void* src, int srctype, int src_be, uint32_t src_channel, uint32_t src_channels,
                             void* dst, int dsttype, int dst_be, uint32_t dst_channel, uint32_t dst_channels,
                             uint32_t nchannels, uint32_t nframes, void* stream) {
  return bbx_transfer_samples_dither_dev(src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel,
                                         dst_channels, nchannels, nframes, BBX_DITHER_NONE, 0, stream);
}

int bbx_transfer_samples_dither_dev(const void* src, int srctype, int src_be, uint32_t src_channel, uint32_t src_channels,
                                    void* dst, int dsttype, int dst_be, uint32_t dst_channel, uint32_t dst_channels,
                                    uint32_t nchannels, uint32_t nframes, int dither, uint64_t seed, void* stream) {
  BBX_REQUIRE(dither == BBX_DITHER_NONE || dither == BBX_DITHER_TPDF, "bbx_transfer_samples: unknown dither mode %d", dither);
  if (!bbx_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 1))
    return BBX_OK;  // silent no-op like the reference
  if (!valid_fmt(srctype) || !valid_fmt(dsttype)) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_transfer_samples_dev: null buffer");
  return transfer_dev_checked(src, srctype, src_be != 0, src_channel, src_channels, dst, dsttype, dst_be != 0, dst_channel,
                              dst_channels, nchannels, nframes, (cudaStream_t)stream, dither, seed);
}

int bbx_transfer_samples(const void* src, int srctype, int src_be, uint32_t src_channel, uint32_t src_channels,
                         void* dst, int dsttype, int dst_be, uint32_t dst_channel, uint32_t dst_channels,
                         uint32_t nchannels, uint32_t nframes) {
  return bbx_transfer_samples_dither(src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel,
                                     dst_channels, nchannels, nframes, BBX_DITHER_NONE, 0);
}

int bbx_transfer_samples_dither(const void* src, int srctype, int src_be, uint32_t src_channel, uint32_t src_channels,
                                void* dst, int dsttype, int dst_be, uint32_t dst_channel, uint32_t dst_channels,
                                uint32_t nchannels, uint32_t nframes, int dither, uint64_t seed) {
  BBX_REQUIRE(dither == BBX_DITHER_NONE || dither == BBX_DITHER_TPDF, "bbx_transfer_samples: unknown dither mode %d", dither);
  if (!bbx_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 1))
    return BBX_OK;
  if (!valid_fmt(srctype) || !valid_fmt(dsttype)) return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_transfer_samples: null buffer");
  int rc = require_device();
  if (rc) return rc;
  uint32_t srclen = fmt_bytes(srctype), dstlen = fmt_bytes(dsttype);
  size_t n = (size_t)nchannels * nframes;
  DeviceScratch &ds = scratch(0), &dd = scratch(1);
  if ((rc = ds.ensure(n * srclen))) return rc;
  if ((rc = dd.ensure(n * dstlen))) return rc;
  cudaStream_t st = cudaStreamPerThread;
  if ((rc = copy_rect_h2d(ds.ptr, (const uint8_t*)src, src_channel, src_channels, nchannels, nframes, srclen, st))) return rc;
  if ((rc = transfer_dev_checked(ds.ptr, srctype, src_be != 0, 0, nchannels, dd.ptr, dsttype, dst_be != 0, 0, nchannels,
                                 nchannels, nframes, st, dither, seed)))
    return rc;
  if ((rc = copy_rect_d2h((uint8_t*)dst, dd.ptr, dst_channel, dst_channels, nchannels, nframes, dstlen, st))) return rc;
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  return BBX_OK;
}

int bbx_transfer_samples_linear(const void* src, int srctype, void* dst, int dsttype, uint32_t nsamples) {
  // src/SoundFormatConversions.cpp:204-219: one frame of nsamples, machine (little) endianness
  if (srctype < 0 || srctype >= FMT_COUNT || dsttype < 0 || dsttype >= FMT_COUNT) return BBX_OK;
  if (!valid_fmt(srctype) || !valid_fmt(dsttype) || nsamples == 0) return BBX_OK;
  return bbx_transfer_samples(src, srctype, 0, 0, nsamples, dst, dsttype, 0, 0, nsamples, nsamples, 1);
}

}  // extern "C"

// ---- MixSamples --------------------------------------------------------------------------
template <typename T>
static int mix_host(const T* src, uint32_t src_channel, uint32_t src_channels, T* dst, uint32_t dst_channel,
                    uint32_t dst_channels, uint32_t nchannels, uint32_t nframes, T mul, const float* gains) {
  int rc = require_device();
  if (rc) return rc;
  size_t n = (size_t)nchannels * nframes;
  DeviceScratch &ds = scratch(0), &dd = scratch(1), &dg = scratch(2);
  if ((rc = ds.ensure(n * sizeof(T)))) return rc;
  if ((rc = dd.ensure(n * sizeof(T)))) return rc;
  cudaStream_t st = cudaStreamPerThread;
  if ((rc = copy_rect_h2d(ds.ptr, (const uint8_t*)src, src_channel, src_channels, nchannels, nframes, sizeof(T), st))) return rc;
  if ((rc = copy_rect_h2d(dd.ptr, (const uint8_t*)dst, dst_channel, dst_channels, nchannels, nframes, sizeof(T), st))) return rc;
  const float* dgains = nullptr;
  if (gains) {
    if ((rc = dg.ensure((size_t)nframes * sizeof(float)))) return rc;
    BBX_CUDA_TRY(cudaMemcpyAsync(dg.ptr, gains, (size_t)nframes * sizeof(float), cudaMemcpyHostToDevice, st));
    dgains = (const float*)dg.ptr;
  }
  uint32_t blocks = (uint32_t)std::min<size_t>((n + 255) / 256, 148u * 16u);
  k_mix<T><<<blocks, 256, 0, st>>>((const T*)ds.ptr, nchannels, (T*)dd.ptr, nchannels, nchannels, nframes, mul, dgains);
  BBX_CUDA_TRY(cudaGetLastError());
  if ((rc = copy_rect_d2h((uint8_t*)dst, dd.ptr, dst_channel, dst_channels, nchannels, nframes, sizeof(T), st))) return rc;
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  return BBX_OK;
}

extern "C" {

int bbx_mix_samples_f32(const float* src, uint32_t src_channel, uint32_t src_channels, float* dst,
                        uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes, float mul) {
  if (!bbx_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 1) ||
      !(mul != 0.0f))
    return BBX_OK;  // (mul != T()) : zero gain is a no-op (SoundMixing.h:65-69)
  BBX_REQUIRE(src && dst, "bbx_mix_samples_f32: null buffer");
  return mix_host<float>(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul, nullptr);
}

int bbx_mix_samples_f64(const double* src, uint32_t src_channel, uint32_t src_channels, double* dst,
                        uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes, double mul) {
  if (!bbx_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 1) ||
      !(mul != 0.0))
    return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_mix_samples_f64: null buffer");
  return mix_host<double>(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul, nullptr);
}

int bbx_interpolator_step(float* st, float inc, uint32_t nsteps) {
  BBX_REQUIRE(st != nullptr, "bbx_interpolator_step: null state");
  for (uint32_t i = 0; i < nsteps; i++) st[1] = interp_step(st[0], st[1], inc);
  return BBX_OK;
}

// per-frame gains g_0 = current, g_{i+1} = step(g_i) (src/SoundMixing.cpp:43-50); advances st
static void ramp_gains(float* st, float inc, uint32_t nframes, std::vector<float>& g) {
  g.resize(nframes);
  for (uint32_t i = 0; i < nframes; i++) {
    g[i] = st[1];
    st[1] = interp_step(st[0], st[1], inc);
  }
}

int bbx_mix_samples_interp(const float* src, uint32_t src_channel, uint32_t src_channels, float* dst,
                           uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes,
                           float* st, float inc) {
  BBX_REQUIRE(st != nullptr, "bbx_mix_samples_interp: null interpolator state");
  // per-frame gain: the contiguous collapse is not allowed (src/SoundMixing.cpp:32-36)
  if (!bbx_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 0) ||
      !((st[1] != 0.0f) || (st[0] != 0.0f)))
    return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_mix_samples_interp: null buffer");
  std::vector<float> g;
  ramp_gains(st, inc, nframes, g);
  return mix_host<float>(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, 1.0f, g.data());
}

int bbx_mix_samples_f32_dev(const float* src, uint32_t src_channel, uint32_t src_channels, float* dst,
                            uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes,
                            float mul, void* stream) {
  if (!bbx_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 1) ||
      !(mul != 0.0f))
    return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_mix_samples_f32_dev: null buffer");
  size_t n = (size_t)nchannels * nframes;
  uint32_t blocks = (uint32_t)std::min<size_t>((n + 255) / 256, 148u * 16u);
  k_mix<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(src + src_channel, src_channels, dst + dst_channel, dst_channels,
                                                         nchannels, nframes, mul, nullptr);
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

int bbx_mix_samples_interp_dev(const float* src, uint32_t src_channel, uint32_t src_channels, float* dst,
                               uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes,
                               float* st, float inc, void* stream) {
  BBX_REQUIRE(st != nullptr, "bbx_mix_samples_interp_dev: null interpolator state");
  if (!bbx_block_transfer_sanity_checks(&src_channel, &src_channels, &dst_channel, &dst_channels, &nchannels, &nframes, 0) ||
      !((st[1] != 0.0f) || (st[0] != 0.0f)))
    return BBX_OK;
  BBX_REQUIRE(src && dst, "bbx_mix_samples_interp_dev: null buffer");
  std::vector<float> g;
  ramp_gains(st, inc, nframes, g);
  DeviceScratch& dg = scratch(2);
  int rc = dg.ensure((size_t)nframes * sizeof(float));
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  // pageable source: the copy is staged before the call returns, so `g` may go out of scope
  BBX_CUDA_TRY(cudaMemcpyAsync(dg.ptr, g.data(), (size_t)nframes * sizeof(float), cudaMemcpyHostToDevice, s));
  size_t n = (size_t)nchannels * nframes;
  uint32_t blocks = (uint32_t)std::min<size_t>((n + 255) / 256, 148u * 16u);
  k_mix<float><<<blocks, 256, 0, s>>>(src + src_channel, src_channels, dst + dst_channel, dst_channels, nchannels, nframes,
                                      1.0f, (const float*)dg.ptr);
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

// ---- FractionalSample --------------------------------------------------------------------
uint32_t bbx_fractional_sample_additional_delay_required(void) { return 14; }  // FractionalSample.cpp:249-252

}  // extern "C"

template <typename T>
static int frac_host(const T* buffer, uint32_t channel, uint32_t channels, uint32_t length, const double* pos, uint32_t n,
                     double* out) {
  BBX_REQUIRE(buffer && pos && out, "bbx_fractional_samples: null buffer");
  BBX_REQUIRE(channels > 0 && channel < channels && length >= 14, "bbx_fractional_samples: bad geometry");
  if (n == 0) return BBX_OK;
  int rc = require_device();
  if (rc) return rc;
  DeviceScratch &db = scratch(0), &dp = scratch(1), &dout = scratch(2);
  size_t bbytes = (size_t)channels * length * sizeof(T);
  if ((rc = db.ensure(bbytes))) return rc;
  if ((rc = dp.ensure((size_t)n * sizeof(double)))) return rc;
  if ((rc = dout.ensure((size_t)n * sizeof(double)))) return rc;
  cudaStream_t st = cudaStreamPerThread;
  BBX_CUDA_TRY(cudaMemcpyAsync(db.ptr, buffer, bbytes, cudaMemcpyHostToDevice, st));
  BBX_CUDA_TRY(cudaMemcpyAsync(dp.ptr, pos, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  k_frac<T><<<ceil_div(n, 256), 256, 0, st>>>((const T*)db.ptr, channel, channels, length, (const double*)dp.ptr, n,
                                              (double*)dout.ptr);
  BBX_CUDA_TRY(cudaGetLastError());
  BBX_CUDA_TRY(cudaMemcpyAsync(out, dout.ptr, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  return BBX_OK;
}

extern "C" {

int bbx_fractional_samples_f32(const float* buffer, uint32_t channel, uint32_t channels, uint32_t length,
                               const double* pos, uint32_t n, double* out) {
  return frac_host<float>(buffer, channel, channels, length, pos, n, out);
}
int bbx_fractional_samples_f64(const double* buffer, uint32_t channel, uint32_t channels, uint32_t length,
                               const double* pos, uint32_t n, double* out) {
  return frac_host<double>(buffer, channel, channels, length, pos, n, out);
}
int bbx_fractional_samples_f32_dev(const float* buffer, uint32_t channel, uint32_t channels, uint32_t length,
                                   const double* pos, uint32_t n, double* out, void* stream) {
  BBX_REQUIRE(buffer && pos && out, "bbx_fractional_samples_f32_dev: null buffer");
  BBX_REQUIRE(channels > 0 && channel < channels && length >= 14, "bbx_fractional_samples_f32_dev: bad geometry");
  if (n == 0) return BBX_OK;
  k_frac<float><<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(buffer, channel, channels, length, pos, n, out);
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

// ---- SoundDelayBuffer --------------------------------------------------------------------
struct bbx_delay {
  uint8_t* buf = nullptr;  // device, interleaved [buflen][channels] in `format`
  int format = FMT_F32;
  uint32_t channels = 0, bytesperframe = 0, buflen = 0, writepos = 0;
  int device = 0;
  // SoundRingBuffer (src/SoundDelayBuffer.h:105-181, src/SoundDelayBuffer.cpp:195-304): the same buffer with a read
  // position that limits writes, reads and both increments; the entry points branch where the reference dispatches
  // virtually
  bool ring = false;
  uint32_t readpos = 0;
};

static uint32_t ring_read_available(const bbx_delay* d) { return d->buflen ? (d->writepos + d->buflen - d->readpos) % d->buflen : 0; }
static uint32_t ring_write_available(const bbx_delay* d) {
  return d->buflen ? (d->readpos + d->buflen - d->writepos - 1) % d->buflen : 0;  // one frame stays free (.h:123)
}

int bbx_delay_create(bbx_delay** out) {
  BBX_REQUIRE(out != nullptr, "bbx_delay_create: null out pointer");
  int rc = require_device();
  if (rc) return rc;
  bbx_delay* d = new bbx_delay();
  cudaError_t ce = cudaGetDevice(&d->device);
  if (ce != cudaSuccess) {
    delete d;
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(ce));
    return BBX_ERR_CUDA;
  }
  *out = d;
  return BBX_OK;
}

int bbx_ring_create(bbx_delay** out) {
  int rc = bbx_delay_create(out);
  if (rc) return rc;
  (*out)->ring = true;
  return BBX_OK;
}
uint32_t bbx_ring_get_read_position(const bbx_delay* d) { return d ? d->readpos : 0; }
uint32_t bbx_ring_get_read_frames_available(const bbx_delay* d) { return d ? ring_read_available(d) : 0; }
uint32_t bbx_ring_get_write_frames_available(const bbx_delay* d) { return d ? ring_write_available(d) : 0; }
int bbx_ring_increment_read_position(bbx_delay* d, uint32_t nframes) {
  // src/SoundDelayBuffer.h:175
  BBX_REQUIRE(d != nullptr && d->ring, "bbx_ring_increment_read_position: not a ring buffer handle");
  if (d->buflen) d->readpos = (d->readpos + std::min(nframes, ring_read_available(d))) % d->buflen;
  return BBX_OK;
}

int bbx_delay_destroy(bbx_delay* d) {
  if (!d) return BBX_OK;
  DeviceGuard dg(d->device);
  if (d->buf) cudaFree(d->buf);
  delete d;
  return BBX_OK;
}

int bbx_delay_set_size(bbx_delay* d, uint32_t chans, uint32_t length, int format) {
  // src/SoundDelayBuffer.cpp:26-61
  BBX_REQUIRE(d != nullptr, "bbx_delay_set_size: null handle");
  BBX_REQUIRE(valid_fmt(format), "bbx_delay_set_size: bad format %d", format);
  chans = std::max(chans, 1u);
  length = std::max(length, 1u);
  if (chans == d->channels && length == d->buflen && format == d->format) return BBX_OK;
  DeviceGuard dg(d->device);
  uint32_t bps = fmt_bytes(format);
  uint8_t* nb = nullptr;
  size_t bytes = (size_t)chans * length * bps;
  BBX_CUDA_TRY(cudaMalloc((void**)&nb, bytes));
  // every failure below releases the new buffer; the object keeps its old ring
  struct Drop {
    uint8_t*& p;
    ~Drop() {
      if (p) cudaFree(p);
    }
  } drop{nb};
  cudaStream_t st = cudaStreamPerThread;
  BBX_CUDA_TRY(cudaMemsetAsync(nb, 0, bytes, st));
  if (d->buf) {
    // the reference maps the old contents across frame by frame (.cpp:46-49).  That is only memory-safe
    // when the format is unchanged and the ring does not shrink; outside that domain (UB there) the copy
    // is clamped to the new length and skipped on a format change.
    if (format == d->format) {
      uint32_t sc = 0, scs = d->channels, dc = 0, dcs = chans, nch = ~0u, nfr = std::min(d->buflen, length);
      if (bbx_block_transfer_sanity_checks(&sc, &scs, &dc, &dcs, &nch, &nfr, 1)) {
        int rc = transfer_dev_checked(d->buf, d->format, false, sc, scs, nb, d->format, false, dc, dcs, nch, nfr, st);
        if (rc) return rc;
      }
    }
    BBX_CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(d->buf);
  }
  BBX_CUDA_TRY(cudaStreamSynchronize(st));
  d->buf = nb;
  nb = nullptr;  // owned by the object now
  d->channels = chans;
  d->buflen = length;
  d->format = format;
  d->writepos %= d->buflen;
  d->bytesperframe = chans * bps;
  if (d->ring) d->readpos %= d->buflen;  // src/SoundDelayBuffer.cpp:214-218
  return BBX_OK;
}

uint32_t bbx_delay_get_channels(const bbx_delay* d) { return d ? d->channels : 0; }
uint32_t bbx_delay_get_length(const bbx_delay* d) { return d ? d->buflen : 0; }
uint32_t bbx_delay_get_write_position(const bbx_delay* d) { return d ? d->writepos : 0; }
int bbx_delay_get_format(const bbx_delay* d) { return d ? d->format : 0; }
const void* bbx_delay_get_buffer_dev(const bbx_delay* d) { return d ? d->buf : nullptr; }

uint32_t bbx_delay_write_samples(bbx_delay* d, const void* vsrc, int srcformat, uint32_t channel, uint32_t nchannels,
                                 uint32_t nframes) {
  // src/SoundDelayBuffer.cpp:77-116; does not move the write position
  if (!d || !d->buf || !vsrc || !valid_fmt(srcformat) || nframes == 0) return 0;
  if (d->ring) {  // SoundRingBuffer::WriteSamples (src/SoundDelayBuffer.cpp:234-254): limited by the space up to the read position
    nframes = std::min(nframes, ring_write_available(d));
    if (nframes == 0) return 0;
  }
  uint32_t srclen = fmt_bytes(srcformat), pos = d->writepos, frames = 0;
  channel = std::min(channel, d->channels - 1);
  nchannels = std::min(nchannels, d->channels - channel);
  if (nchannels == 0) return 0;
  DeviceGuard dg(d->device);
  DeviceScratch& ds = scratch(0);
  size_t bytes = (size_t)nchannels * srclen * nframes;
  if (ds.ensure(bytes)) return 0;
  cudaStream_t st = cudaStreamPerThread;
  if (cudaMemcpyAsync(ds.ptr, vsrc, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return 0;
  const uint8_t* src = (const uint8_t*)ds.ptr;
  while (nframes) {
    uint8_t* dst = d->buf + (size_t)pos * d->bytesperframe;
    uint32_t n = std::min(nframes, d->buflen - pos);
    uint32_t sc = 0, scs = nchannels, dc = channel, dcs = d->channels, nch = nchannels, nfr = n;
    if (bbx_block_transfer_sanity_checks(&sc, &scs, &dc, &dcs, &nch, &nfr, 1))
      if (transfer_dev_checked(src, srcformat, false, sc, scs, dst, d->format, false, dc, dcs, nch, nfr, st)) return frames;
    src += (size_t)nchannels * srclen * n;
    pos = (pos + n) % d->buflen;
    nframes -= n;
    frames += n;
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) return 0;
  return frames;
}

int bbx_delay_increment_write_position(bbx_delay* d, uint32_t nframes) {
  BBX_REQUIRE(d != nullptr, "bbx_delay_increment_write_position: null handle");
  if (d->ring) nframes = std::min(nframes, ring_write_available(d));  // src/SoundDelayBuffer.h:148
  if (d->buflen) d->writepos = (d->writepos + nframes) % d->buflen;
  return BBX_OK;
}

uint32_t bbx_delay_read_samples(bbx_delay* d, void* vdst, int dstformat, uint32_t delay, uint32_t channel,
                                uint32_t nchannels, uint32_t nframes) {
  // src/SoundDelayBuffer.cpp:134-170
  if (!d || !d->buf || !vdst || !valid_fmt(dstformat)) return 0;
  if (d->ring) {
    // SoundRingBuffer::ReadSamples (src/SoundDelayBuffer.cpp:274-303): delay limited to (read - write) mod length, the frame
    // count to (write + delay - read) mod length; the base class then reads relative to the WRITE position
    delay = std::min(delay, (d->readpos + d->buflen - d->writepos) % d->buflen);
    nframes = std::min(nframes, (d->writepos + d->buflen + delay - d->readpos) % d->buflen);
  }
  uint32_t dstlen = fmt_bytes(dstformat), frames = 0;
  uint32_t pos = (d->writepos + d->buflen - delay) % d->buflen;
  channel = std::min(channel, d->channels - 1);
  nchannels = std::min(nchannels, d->channels - channel);
  nframes = std::min(nframes, delay);  // cannot read past the write position
  if (nchannels == 0 || nframes == 0) return 0;
  DeviceGuard dg(d->device);
  DeviceScratch& dd = scratch(1);
  size_t bytes = (size_t)nchannels * dstlen * nframes;
  if (dd.ensure(bytes)) return 0;
  cudaStream_t st = cudaStreamPerThread;
  uint8_t* dst = (uint8_t*)dd.ptr;
  uint32_t left = nframes;
  while (left) {
    const uint8_t* src = d->buf + (size_t)pos * d->bytesperframe;
    uint32_t n = std::min(left, d->buflen - pos);
    uint32_t sc = channel, scs = d->channels, dc = 0, dcs = nchannels, nch = nchannels, nfr = n;
    if (bbx_block_transfer_sanity_checks(&sc, &scs, &dc, &dcs, &nch, &nfr, 1))
      if (transfer_dev_checked(src, d->format, false, sc, scs, dst, dstformat, false, dc, dcs, nch, nfr, st)) return frames;
    dst += (size_t)nchannels * dstlen * n;
    pos = (pos + n) % d->buflen;
    left -= n;
    frames += n;
  }
  if (cudaMemcpyAsync(vdst, dd.ptr, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return 0;
  if (cudaStreamSynchronize(st) != cudaSuccess) return 0;
  return frames;
}

float bbx_delay_read_sample(bbx_delay* d, uint32_t channel, uint32_t delay) {
  // src/SoundDelayBuffer.cpp:176-191.  The reference adds `channel` as a BYTE offset (.cpp:187), which is
  // only right for channel 0; here the offset is in samples (documented deviation for channel > 0).
  float res = 0.0f;
  if (!d || !d->buf || channel >= d->channels) return res;
  uint32_t pos = (d->writepos + d->buflen - delay) % d->buflen;
  DeviceGuard dg(d->device);
  DeviceScratch& dd = scratch(1);
  if (dd.ensure(sizeof(float))) return res;
  cudaStream_t st = cudaStreamPerThread;
  if (transfer_dev_checked(d->buf + (size_t)pos * d->bytesperframe, d->format, false, channel, d->channels, dd.ptr, FMT_F32,
                           false, 0, 1, 1, 1, st))
    return res;
  cudaMemcpyAsync(&res, dd.ptr, sizeof(float), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  return res;
}

uint32_t bbx_delay_copy_buffer(const bbx_delay* d, void* dst, uint32_t maxbytes) {
  if (!d || !d->buf || !dst) return 0;
  uint32_t bytes = d->bytesperframe * d->buflen;
  if (bytes > maxbytes) return 0;
  DeviceGuard dg(d->device);
  if (cudaMemcpy(dst, d->buf, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return bytes;
}

}  // extern "C"
