// kernels_pcm.cuh -- both PCM ends of the engine: interleaved PCM in any SampleFormat_t -> planar fp32 blocks (k_pcm_in*), and
// delay ring -> delayed reads, delay crossfade, MixSamples-order mixdown -> PCM (k_pcm_out*).  Instantiated per (format, access).
#pragma once

#include "kernels_common.cuh"
#include "formats.cuh"
#include "fracsample.cuh"

namespace bbx {

// ------------------------------------------------------------------------------------------
// k_pcm_in
// ------------------------------------------------------------------------------------------
struct PcmInArgs {
  const uint8_t* pcm;
  int fmt;
  int be;
  uint32_t in_channels, n_inputs;
  uint32_t B, T;
  float* xin_cur;         // [n_inputs][xstride]
  const float* xin_prev;  // previous call's buffer
  uint32_t xstride;
  uint32_t prev_off;      // offset of the previous call's last block inside xin_prev rows
  int fast;               // little-endian, base and frame stride aligned to the sample size: typed loads
};

template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_in(PcmInArgs a) {
  __shared__ float tile[32][33];
  const uint32_t f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool is_prev = f0 < a.B;  // B is a multiple of 32: a tile never straddles the boundary
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  if (!is_prev) {
    // phase 1: lanes over channels (contiguous bytes within a frame)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t fl = warp + 8 * i, c = c0 + lane;
      uint32_t frame = f0 + fl - a.B;
      float v = 0.f;
      if (c < a.n_inputs) v = load_as_f32_t<FMT, ACC>(a.pcm + ((uint64_t)frame * a.in_channels + c) * bps);
      tile[fl][lane] = v;
    }
    __syncthreads();
  }
  // phase 2: lanes over frames (contiguous floats of one planar row)
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t cl = warp + 8 * i, c = c0 + cl;
    if (c >= a.n_inputs) continue;
    uint32_t f = f0 + lane;
    float v = is_prev ? a.xin_prev[(uint64_t)c * a.xstride + a.prev_off + f] : tile[lane][cl];
    a.xin_cur[(uint64_t)c * a.xstride + f] = v;
  }
}

// Same transpose with 128-frame tiles (B % 128 == 0): 16 independent loads per thread are in flight before the first
// shared-memory store (the 32-frame kernel above is bound by the latency of its 4), a quarter of the CTAs.
template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_in128(PcmInArgs a) {
  __shared__ float tile[128][33];
  const uint32_t f0 = blockIdx.x * 128, c0 = blockIdx.y * 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool is_prev = f0 < a.B;  // B is a multiple of 128: a tile never straddles the boundary
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  if (!is_prev) {
    // phase 1: lanes over channels (contiguous bytes within a frame)
    float v[16];
    const uint32_t c = c0 + lane;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t frame = f0 + warp + 8 * i - a.B;
      v[i] = (c < a.n_inputs) ? load_as_f32_t<FMT, ACC>(a.pcm + ((uint64_t)frame * a.in_channels + c) * bps) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) tile[warp + 8 * i][lane] = v[i];
    __syncthreads();
  }
  // phase 2: lanes over frames (contiguous floats of one planar row); thread: channels warp + {0, 8, 16, 24}, 4 x 32 frames
  float o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t cl = warp + 8 * (i >> 2), c = c0 + cl, fl = (i & 3) * 32 + lane;
    o[i] = 0.f;
    if (c < a.n_inputs) o[i] = is_prev ? a.xin_prev[(uint64_t)c * a.xstride + a.prev_off + f0 + fl] : tile[fl][cl];
  }
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t cl = warp + 8 * (i >> 2), c = c0 + cl, fl = (i & 3) * 32 + lane;
    if (c < a.n_inputs) a.xin_cur[(uint64_t)c * a.xstride + f0 + fl] = o[i];
  }
}

// ------------------------------------------------------------------------------------------
// k_pcm_out : delay read + mixdown + format conversion
// ------------------------------------------------------------------------------------------
// one routed path in mixdown order (ascending stream per output == MixSamples call order), everything k_pcm_out
// needs about it in one place
struct RouteEntry {
  uint32_t stream;  // delay ring of the path
  float gain;
  uint32_t icur, iold;  // (uint32)delay mod Rd for the integer-delay mode: in force after / before this call's first block boundary
  uint32_t flags;       // bit0: crossfade old->cur over the first block
  uint32_t pad;
  double dcur, dold;    // the same delays in samples (fractional mode)
};
static_assert(sizeof(RouteEntry) == 40, "RouteEntry layout");

struct RouteView {
  const uint32_t* out_first;  // [n_outputs+1] CSR over outputs
  const RouteEntry* entry;    // per route
};

struct PcmOutArgs {
  uint8_t* pcm;
  int fmt;
  int be;
  uint32_t out_channels, n_outputs;
  uint32_t B, T;
  const float* ybuf;
  uint32_t Rd, wpos0;
  int fractional;
  int fast;  // typed stores (see PcmInArgs::fast)
  RouteView rv;
};

__device__ __forceinline__ float delayed_read(const float* __restrict__ ring, uint32_t Rd, uint32_t w, uint32_t n, double d,
                                              uint32_t di, int fractional) {
  if (fractional) {
    // FractionalSample(ring, 0, 1, Rd, fmod((w + n + Rd) - d, Rd))   (src/FractionalSample.cpp:312-341)
    const double pos = fmod((double)(w + n + Rd) - d, (double)Rd);
    return __double2float_rn(fractional_sample_dev<float>(ring, 0, 1, Rd, pos));
  }
  // ring[(w + n - d) mod R] with d = (uint)delay mod R precomputed on the host   (src/SoundDelayBuffer.cpp:141)
  uint32_t idx = w + n + Rd - di;  // w < Rd, n < B <= Rd, di < Rd  ->  idx < 3 Rd
  if (idx >= Rd) idx -= Rd;
  if (idx >= Rd) idx -= Rd;
  return ring[idx];
}

static constexpr uint32_t kPcmOutCache = 64;  // routes of one 32-output tile kept in shared memory

template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_out(PcmOutArgs a) {
  __shared__ float tile[32][33];
  __shared__ uint32_t s_first[33];
  __shared__ RouteEntry s_rt[kPcmOutCache];
  const uint32_t f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t t = f0 / a.B;               // a tile lies inside one block (B % 32 == 0)
  const uint32_t w = (a.wpos0 + t * a.B) % a.Rd;
  const float inc = 1.0f / (float)a.B;
  // the tile's slice of the route tables -> shared memory (two dependent loads per CTA instead of a chain of table
  // lookups per sample); tiles with more than kPcmOutCache routes read the entries from global memory
  const uint32_t no = min(32u, a.n_outputs - c0);
  if (threadIdx.x <= no) s_first[threadIdx.x] = a.rv.out_first[c0 + threadIdx.x];
  __syncthreads();
  const uint32_t r0 = s_first[0], nr = s_first[no] - r0;
  const bool cached = nr <= kPcmOutCache;
  if (cached && threadIdx.x < nr) s_rt[threadIdx.x] = a.rv.entry[r0 + threadIdx.x];
  __syncthreads();
  // phase 1: lanes over frames (ring reads are contiguous), one output channel per warp pass
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    const uint32_t cl = warp + 8 * i, o = c0 + cl;
    float bus = 0.f;
    if (o < a.n_outputs) {
      const uint32_t n = f0 + lane - t * a.B;  // frame inside the block
      const uint32_t rb = s_first[cl], re = s_first[cl + 1];
      for (uint32_t r = rb; r < re; r++) {  // ascending stream order == MixSamples call order
        const RouteEntry en = cached ? s_rt[r - r0] : a.rv.entry[r];
        if (!(en.gain != 0.0f)) continue;  // (mul != T()): a zero gain is a no-op (src/SoundMixing.h:65-69)
        const float* ring = a.ybuf + (uint64_t)en.stream * a.Rd;
        float v = delayed_read(ring, a.Rd, w, n, en.dcur, en.icur, a.fractional);
        if (t == 0 && (en.flags & 1u)) {
          const float vo = delayed_read(ring, a.Rd, w, n, en.dold, en.iold, a.fractional);
          const float g = __fmul_rn((float)n, inc);
          v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, g), vo), __fmul_rn(g, v));
        }
        bus = __fadd_rn(bus, __fmul_rn(en.gain, v));  // dst += mul * src, rounded separately
      }
    }
    tile[lane][cl] = bus;
  }
  __syncthreads();
  // phase 2: lanes over channels (contiguous bytes of one interleaved frame)
  constexpr uint32_t bps = FmtBytes<FMT>::value;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t fl = warp + 8 * i, o = c0 + lane;
    if (o >= a.n_outputs) continue;
    const uint32_t frame = f0 + fl;
    store_from_f32_t<FMT, ACC>(a.pcm + ((uint64_t)frame * a.out_channels + o) * bps, tile[fl][lane]);
  }
}

// Fractional-delay engines (14-tap polyphase reads in double, C4): one sample per thread.  The tile kernel above gives a
// thread four samples one after the other, each a chain of table -> ring -> 14 dependent double adds; with 16 CTAs for
// a 512-frame block the chains, not the arithmetic, set the kernel time (ncu: 13 % issue, long-scoreboard bound).  Here a
// warp is one frame and a lane one output channel, so the 32 stores of a warp are one contiguous piece of the interleaved
// frame (no shared-memory transpose) and four times as many threads are in flight.  Same per-sample arithmetic and route
// order as k_pcm_out.
template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_out_frac(PcmOutArgs a) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t frame = blockIdx.x * 8 + warp, o = blockIdx.y * 32 + lane;
  if (o >= a.n_outputs) return;
  const uint32_t t = frame / a.B, n = frame - t * a.B;
  const uint32_t w = (a.wpos0 + t * a.B) % a.Rd;
  const float inc = 1.0f / (float)a.B;
  const uint32_t rb = a.rv.out_first[o], re = a.rv.out_first[o + 1];
  float bus = 0.f;
  for (uint32_t r = rb; r < re; r++) {  // ascending stream order == MixSamples call order
    const RouteEntry en = a.rv.entry[r];
    if (!(en.gain != 0.0f)) continue;  // (mul != T()): a zero gain is a no-op (src/SoundMixing.h:65-69)
    const float* ring = a.ybuf + (uint64_t)en.stream * a.Rd;
    float v = delayed_read(ring, a.Rd, w, n, en.dcur, en.icur, a.fractional);
    if (t == 0 && (en.flags & 1u)) {
      const float vo = delayed_read(ring, a.Rd, w, n, en.dold, en.iold, a.fractional);
      const float g = __fmul_rn((float)n, inc);
      v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, g), vo), __fmul_rn(g, v));
    }
    bus = __fadd_rn(bus, __fmul_rn(en.gain, v));  // dst += mul * src, rounded separately
  }
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  store_from_f32_t<FMT, ACC>(a.pcm + ((uint64_t)frame * a.out_channels + o) * bps, bus);
}

// Mixdown of many paths into few outputs (the binaural renderer: 64 sources x 2 ears -> 2 outputs).  The kernel above
// walks the routes of an output one after the other inside one thread: 64 dependent table + ring reads per sample, and
// only n_outputs of its 32 channel slots do anything.  Here every thread of the CTA takes (route, frame) items: the
// delayed reads, the delay crossfade and the products gain * v of ALL routes of a 32-frame tile are formed in parallel
// into shared memory, then one thread per (output, frame) adds the products in ascending route order -- the same
// dst += mul * src with separately rounded product and sum (src/SoundMixing.h:76-79), zero gains skipped, so the bytes are
// those of k_pcm_out (tests: routed engines run both kernels' shapes against the oracle and each other).
static constexpr uint32_t kMixMaxRoutes = 256;

template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_out_mix(PcmOutArgs a) {
  __shared__ float prod[kMixMaxRoutes][32];
  __shared__ RouteEntry s_rt[kMixMaxRoutes];
  __shared__ uint32_t s_first[33];
  const uint32_t f0 = blockIdx.x * 32;
  const uint32_t t = f0 / a.B;  // a tile lies inside one block (B % 32 == 0)
  const uint32_t w = (a.wpos0 + t * a.B) % a.Rd;
  const float inc = 1.0f / (float)a.B;
  const uint32_t no = a.n_outputs;  // <= 32 (host)
  if (threadIdx.x <= no) s_first[threadIdx.x] = a.rv.out_first[threadIdx.x];
  __syncthreads();
  const uint32_t r0 = s_first[0], nr = s_first[no] - r0;  // <= kMixMaxRoutes (host)
  for (uint32_t r = threadIdx.x; r < nr; r += 256) s_rt[r] = a.rv.entry[r0 + r];
  __syncthreads();
  const uint32_t nb0 = f0 - t * a.B;  // frame of the tile's first sample inside its block
  // stage 1: items (route, frame), four per thread and pass so that their ring reads are in flight together
  for (uint32_t base = threadIdx.x; base < nr * 32; base += 4 * 256) {
    float p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t idx = base + 256 * k;
      p[k] = 0.f;
      if (idx < nr * 32) {
        const uint32_t r = idx >> 5, n = nb0 + (idx & 31);
        const RouteEntry en = s_rt[r];
        if (en.gain != 0.0f) {
          const float* ring = a.ybuf + (uint64_t)en.stream * a.Rd;
          float v = delayed_read(ring, a.Rd, w, n, en.dcur, en.icur, a.fractional);
          if (t == 0 && (en.flags & 1u)) {
            const float vo = delayed_read(ring, a.Rd, w, n, en.dold, en.iold, a.fractional);
            const float g = __fmul_rn((float)n, inc);
            v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, g), vo), __fmul_rn(g, v));
          }
          p[k] = __fmul_rn(en.gain, v);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t idx = base + 256 * k;
      if (idx < nr * 32) prod[idx >> 5][idx & 31] = p[k];
    }
  }
  __syncthreads();
  // stage 2: one thread per (output, frame): the ordered sum, then the sample in the output format
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  for (uint32_t item = threadIdx.x; item < no * 32; item += 256) {
    const uint32_t o = item >> 5, fl = item & 31;
    float bus = 0.f;
    for (uint32_t r = s_first[o] - r0; r < s_first[o + 1] - r0; r++)
      if (s_rt[r].gain != 0.0f) bus = __fadd_rn(bus, prod[r][fl]);  // a zero gain is a no-op, not "+ 0"
    store_from_f32_t<FMT, ACC>(a.pcm + ((uint64_t)(f0 + fl) * a.out_channels + o) * bps, bus);
  }
}

// 128-frame tiles (B % 128 == 0).  Outputs fed by exactly one path with an integer delay and no delay crossfade in this
// block (every output of the PER_CHANNEL and MIMO modes in the steady state) issue their four ring reads together; the
// arithmetic is the same dst += mul * src, rounded separately.
template <int FMT, int ACC>
__global__ void __launch_bounds__(256) k_pcm_out128(PcmOutArgs a) {
  __shared__ float tile[128][33];
  __shared__ uint32_t s_first[33];
  __shared__ RouteEntry s_rt[kPcmOutCache];
  const uint32_t f0 = blockIdx.x * 128, c0 = blockIdx.y * 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t t = f0 / a.B;               // a tile lies inside one block (B % 128 == 0)
  const uint32_t w = (a.wpos0 + t * a.B) % a.Rd;
  const float inc = 1.0f / (float)a.B;
  const uint32_t no = min(32u, a.n_outputs - c0);
  if (threadIdx.x <= no) s_first[threadIdx.x] = a.rv.out_first[c0 + threadIdx.x];
  __syncthreads();
  const uint32_t r0 = s_first[0], nr = s_first[no] - r0;
  const bool cached = nr <= kPcmOutCache;
  if (cached && threadIdx.x < nr) s_rt[threadIdx.x] = a.rv.entry[r0 + threadIdx.x];
  __syncthreads();
  // phase 1: lanes over frames; thread: outputs warp + {0, 8, 16, 24}, 4 x 32 frames each.  The route entries of the
  // four outputs are resolved first, then the ring reads of every single-path output are issued together (up to 16
  // loads in flight per thread: with one output after the other the kernel was a chain of four DRAM round trips per
  // CTA behind the two of the route tables), then the general outputs take the per-route loop.
  const uint32_t nb = f0 - t * a.B + lane;  // frame inside the block of the first of the four chunks
  uint32_t q_rb[4], q_re[4], q_stream[4], q_icur[4];
  float q_gain[4];
  bool q_simple[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t cl = warp + 8 * q;
    q_simple[q] = false;
    q_rb[q] = q_re[q] = 0;
    q_stream[q] = q_icur[q] = 0;
    q_gain[q] = 0.f;
    if (c0 + cl < a.n_outputs) {
      q_rb[q] = s_first[cl];
      q_re[q] = s_first[cl + 1];
      if (q_re[q] == q_rb[q] + 1 && !a.fractional) {
        const RouteEntry en = cached ? s_rt[q_rb[q] - r0] : a.rv.entry[q_rb[q]];
        if (!(t == 0 && (en.flags & 1u))) {
          q_simple[q] = true;
          q_stream[q] = en.stream;
          q_icur[q] = en.icur;
          q_gain[q] = en.gain;
        }
      }
    }
  }
  float v[4][4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float* ring = a.ybuf + (uint64_t)q_stream[q] * a.Rd;
#pragma unroll
    for (int k = 0; k < 4; k++)
      v[q][k] = (q_simple[q] && q_gain[q] != 0.0f) ? delayed_read(ring, a.Rd, w, nb + 32 * k, 0.0, q_icur[q], 0) : 0.f;
  }
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t cl = warp + 8 * q;
    float bus[4] = {0.f, 0.f, 0.f, 0.f};
    if (q_simple[q]) {
      if (q_gain[q] != 0.0f) {  // (mul != T()): a zero gain is a no-op (src/SoundMixing.h:65-69)
#pragma unroll
        for (int k = 0; k < 4; k++) bus[k] = __fadd_rn(0.f, __fmul_rn(q_gain[q], v[q][k]));
      }
    } else if (c0 + cl < a.n_outputs) {
      const uint32_t rb = q_rb[q], re = q_re[q];
#pragma unroll 1
      for (int k = 0; k < 4; k++) {
        const uint32_t n = nb + 32 * k;
        float b = 0.f;
        for (uint32_t r = rb; r < re; r++) {  // ascending stream order == MixSamples call order
          const RouteEntry en = cached ? s_rt[r - r0] : a.rv.entry[r];
          if (!(en.gain != 0.0f)) continue;  // (mul != T()): a zero gain is a no-op (src/SoundMixing.h:65-69)
          const float* ring = a.ybuf + (uint64_t)en.stream * a.Rd;
          float vv = delayed_read(ring, a.Rd, w, n, en.dcur, en.icur, a.fractional);
          if (t == 0 && (en.flags & 1u)) {
            const float vo = delayed_read(ring, a.Rd, w, n, en.dold, en.iold, a.fractional);
            const float g = __fmul_rn((float)n, inc);
            vv = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, g), vo), __fmul_rn(g, vv));
          }
          b = __fadd_rn(b, __fmul_rn(en.gain, vv));  // dst += mul * src, rounded separately
        }
        bus[k] = b;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) tile[32 * k + lane][cl] = bus[k];
  }
  __syncthreads();
  // phase 2: lanes over channels (contiguous bytes of one interleaved frame)
  constexpr uint32_t bps = FmtBytes<FMT>::value;
  const uint32_t o = c0 + lane;
  if (o < a.n_outputs) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t fl = warp + 8 * i;
      store_from_f32_t<FMT, ACC>(a.pcm + ((uint64_t)(f0 + fl) * a.out_channels + o) * bps, tile[fl][lane]);
    }
  }
}

}  // namespace bbx
