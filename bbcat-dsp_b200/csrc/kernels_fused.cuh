// kernels_fused.cuh -- k_block_fused: one launch for a whole streaming call (T = 1) of a PER_CHANNEL or ROUTED engine
// with short filters (the latency path of BASELINE.json's configs C1, C2, C4).
//
// The multi-kernel path runs k_pcm_in -> k_rfft -> k_fdl_mac -> k_irfft -> k_pcm_out for such a call: five dependent
// launches of 2 to 64 CTAs with a few microseconds of work each (profiles/r01: 43 us p50 for 0.05 us of roofline work on
// C1).  Here one CTA group per stream (path) does all of it for its stream:
//     PCM (any SampleFormat_t) + the previous block  ->  window  ->  forward transform  ->  FDL slot (HBM, state)
//     -> MAC over the stream's plan segments (H and FDL rows come from L2)  ->  slot sums in plan order
//     -> inverse transform -> overlap-save -> filter crossfade -> delay ring (HBM, state)
//     -> [PER_CHANNEL: delayed read (integer / 14-tap fractional), delay crossfade, gain, output format -> PCM]
// ROUTED engines mix several streams into one output, so their output stage stays a second launch (k_pcm_out*).
// Streams that share an input (ROUTED fan-out) each transform it again; they write identical bytes to the input's FDL
// slot and history row, which is harmless.
//
// Bit-identity with the multi-kernel path is part of the contract (tests/test_fused_gpu.py): the transforms are the
// same device functions called the same way, the MAC walks the stream's segments in plan order with the same four FMAs
// per (row, bin), accumulates per partial-sum slot exactly where k_fdl_mac starts and ends its accumulators, and adds
// the slots (and the Nyquist sums) in slot order like k_irfft does; the output stage is k_pcm_out's arithmetic.
#pragma once

#include "kernels_fft.cuh"
#include "kernels_pcm.cuh"
#include "mac_common.cuh"

namespace bbx {

struct FusedArgs {
  // PCM in / out
  const uint8_t* pcm_in;
  uint8_t* pcm_out;
  int infmt, in_be, in_fast, outfmt, out_be, out_fast;
  int planar_in;  // 1: k_pcm_in ran first, this block is already in the history row (wide interleaved PCM in host memory)
  uint32_t in_channels, out_channels;
  // streams and their inputs
  uint32_t n_streams;
  const uint32_t* stream_input;  // NULL: stream c reads input c (PER_CHANNEL)
  // input history (k_pcm_in's state)
  float* xin_cur;
  const float* xin_prev;
  uint32_t xstride, prev_off;
  // FDL and plan
  float2* fdl;
  uint32_t R, head;
  const float2* tw;
  const MacSeg* segs;
  const uint32_t* job_seg_first;  // [n_jobs + 1]
  const uint32_t* xjob;           // per stream: extra job to crossfade into / kNoJob / kSameJob; NULL: no crossfade this call
  // delay ring and output stage
  float* ybuf;
  uint32_t Rd, wpos;
  int fractional;
  const RouteEntry* entry;  // PER_CHANNEL: entry[c] is the route of output c
};

// MAC of one job for the bins k = tid + h NT of this thread: per-slot accumulators, slots added in plan order.
template <int M>
__device__ __forceinline__ void fused_job_mac(const FusedArgs& a, uint32_t job, int tid, float2 (&total)[FftCfg<M>::R]) {
  constexpr int RAD = FftCfg<M>::R, NT = FftCfg<M>::NT;
  float2 acc[RAD];
  float nacc = 0.f, ntotal = 0.f;
#pragma unroll
  for (int h = 0; h < RAD; h++) total[h] = acc[h] = make_float2(0.f, 0.f);
  const uint32_t s0 = a.job_seg_first[job], s1 = a.job_seg_first[job + 1];
  for (uint32_t si = s0; si < s1; si++) {
    const MacSeg sg = a.segs[si];
    if (sg.flags & 1u) {
#pragma unroll
      for (int h = 0; h < RAD; h++) acc[h] = make_float2(0.f, 0.f);
      nacc = 0.f;
    }
    const float2* hrow = reinterpret_cast<const float2*>(sg.H) + (uint64_t)sg.p0 * M + tid;
    const float2* xch = a.fdl + (uint64_t)sg.fdl_ch * a.R * M + tid;
    int slot = (int)a.head - (int)(sg.p0 % a.R);
    if (slot < 0) slot += (int)a.R;
    // rows in batches of four: the loads of a batch are all in flight before its FMAs (the FMAs stay in row order)
    constexpr int UB = (RAD == 8) ? 2 : 4;
    for (uint32_t p = 0; p < sg.np; p += UB) {
      float2 hv[UB][RAD], xv[UB][RAD];
#pragma unroll
      for (int u = 0; u < UB; u++) {
        const bool on = p + u < sg.np;
        int sl = slot - u;
        if (sl < 0) sl += (int)a.R;
        const float2* xrow = xch + (uint64_t)sl * M;
#pragma unroll
        for (int h = 0; h < RAD; h++) {
          hv[u][h] = on ? __ldg(hrow + (uint64_t)u * M + h * NT) : make_float2(0.f, 0.f);
          xv[u][h] = on ? xrow[h * NT] : make_float2(0.f, 0.f);  // plain load: this thread wrote the newest row a moment ago
        }
      }
#pragma unroll
      for (int u = 0; u < UB; u++)
        if (p + u < sg.np) {
#pragma unroll
          for (int h = 0; h < RAD; h++) cmac(acc[h].x, acc[h].y, hv[u][h].x, hv[u][h].y, xv[u][h].x, xv[u][h].y);
          nacc = fmaf(hv[u][0].y, xv[u][0].y, nacc);  // meaningful in the thread that owns bin 0 only
        }
      hrow += (uint64_t)UB * M;
      slot -= UB;
      if (slot < 0) slot += (int)a.R;
    }
    if (sg.flags & 2u) {
#pragma unroll
      for (int h = 0; h < RAD; h++) {
        total[h].x += acc[h].x;
        total[h].y += acc[h].y;
      }
      ntotal += nacc;
    }
  }
  // bin 0: the complex MAC left G = DC - N in the real part; restore (DC, Nyquist) like k_irfft
  if (tid == 0) total[0] = make_float2(total[0].x + ntotal, ntotal);
}

template <int M, bool FUSE_OUT>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB) k_block_fused(const FusedArgs a) {
  constexpr int RAD = FftCfg<M>::R, NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP;
  constexpr bool R8 = RAD == 8;
  extern __shared__ float2 k_fused_smem[];  // per transform: spectrum x[M] + padded FFT workspace s[MP]
  float2* x = k_fused_smem + (size_t)threadIdx.y * (M + MP);
  float2* s = x + M;
  const int tid = threadIdx.x;
  const uint32_t sq = blockIdx.x * FPB + threadIdx.y;
  const bool active = sq < a.n_streams;
  const uint32_t stream = active ? sq : a.n_streams - 1;  // idle transforms of the last CTA redo a valid one, stores masked
  const uint32_t input = a.stream_input ? a.stream_input[stream] : stream;
  Tw8<R8 ? M : 64> tw8;
  if constexpr (R8) load_tw8<M>(tw8, a.tw, tid);

  // ---- 1. window [previous block | this block] as z[n] = x[2n] + i x[2n+1]; this block also goes to the history row ----
  {
    const uint32_t ibps = fmt_bytes(a.infmt);
    const float2* prev = reinterpret_cast<const float2*>(a.xin_prev + (uint64_t)input * a.xstride + a.prev_off);
    float2* cur = reinterpret_cast<float2*>(a.xin_cur + (uint64_t)input * a.xstride + M);
    float2 v[RAD];
#pragma unroll
    for (int r = 0; r < RAD; r++) {
      const uint32_t n = tid + r * NT;  // z index, 0 .. M-1; the first M/2 come from the previous block
      if (n < (uint32_t)M / 2) {
        v[r] = prev[n];
      } else if (a.planar_in) {
        v[r] = cur[n - M / 2];
      } else {
        const uint32_t f = 2 * n - M;  // frame inside this block
        const uint8_t* p = a.pcm_in + ((uint64_t)f * a.in_channels + input) * ibps;
        v[r].x = load_as_f32(p, a.infmt, a.in_be != 0, a.in_fast != 0);
        v[r].y = load_as_f32(p + (uint64_t)a.in_channels * ibps, a.infmt, a.in_be != 0, a.in_fast != 0);
        if (active) cur[n - M / 2] = v[r];
      }
    }
    // ---- 2. forward transform -> FDL slot ----
    float2* row = a.fdl + ((uint64_t)input * a.R + a.head) * M;
    if constexpr (R8) {
      fft_bar<M>();
      pass8_first<M, false>(v, s, tid);
      passes8_rest<M, false>(s, tw8, tid);
      rfft_split_store8<M>(s, tw8, row, 1.0f, tid, active);
    } else {
#pragma unroll
      for (int r = 0; r < RAD; r++) s[PADM<M>(tid + r * NT)] = v[r];
      __syncthreads();
      cfft_smem<M, false>(s, a.tw, tid);
      rfft_split_store<M>(s, a.tw, row, 1.0f, tid, active);
    }
  }
  // every bin this thread reads back below (k = tid + h NT) it has just written itself; streams sharing the input wrote
  // the same bytes.  The barrier keeps the workspace reuse of the inverse transform behind the split stage.
  if constexpr (R8) fft_bar<M>();
  else __syncthreads();

  // ---- 3./4. MAC -> inverse transform -> overlap-save (-> the same for the crossfade partner) ----
  float o[RAD], o2[RAD];
  auto job_to_time = [&](uint32_t job, bool mine, float (&out)[RAD]) {
    float2 tot[RAD];
    if (mine) fused_job_mac<M>(a, job, tid, tot);
    else {
#pragma unroll
      for (int h = 0; h < RAD; h++) tot[h] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < RAD; h++) x[tid + h * NT] = tot[h];
    if constexpr (R8) {
      fft_bar<M>();
      float2 v[8];
      irfft_unsplit8<M>(x, tw8, v, tid);
      pass8_first<M, true>(v, s, tid);
      passes8_rest<M, true>(s, tw8, tid);
#pragma unroll
      for (int h = 0; h < 4; h++) {
        const float2 z = s[PADM<M>(M / 2 + tid + h * NT)];
        out[2 * h] = z.x;
        out[2 * h + 1] = z.y;
      }
    } else {
      __syncthreads();
      irfft_unsplit<M>(x, a.tw, s, tid);
      __syncthreads();
      cfft_smem<M, true>(s, a.tw, tid);
#pragma unroll
      for (int h = 0; h < RAD / 2; h++) {
        const float2 z = s[PADM<M>(M / 2 + tid + h * NT)];
        out[2 * h] = z.x;
        out[2 * h + 1] = z.y;
      }
      __syncthreads();
    }
  };
  job_to_time(stream, true, o);
  const uint32_t xj = a.xjob ? a.xjob[stream] : kNoJob;
  // block-wide barriers inside the transforms: every transform of the CTA runs the second pass when any needs it
  const int any_x = __syncthreads_or(xj != kNoJob && xj != kSameJob);
  if (any_x) {
    const bool mine = (xj != kNoJob && xj != kSameJob);
    job_to_time(mine ? xj : 0u, mine, o2);
  }
  if (xj != kNoJob) {
    if (xj == kSameJob) {
#pragma unroll
      for (int i = 0; i < RAD; i++) o2[i] = o[i];
    }
    const float inc = 1.0f / (float)M;
#pragma unroll
    for (int h = 0; h < RAD / 2; h++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const uint32_t n = 2 * (tid + h * NT) + c;
        const float g = __fmul_rn((float)n, inc);
        const float va = __fmul_rn(__fsub_rn(1.0f, g), o[2 * h + c]);
        const float vb = __fmul_rn(g, o2[2 * h + c]);
        o[2 * h + c] = __fadd_rn(va, vb);
      }
  }
  // ---- 5. delay ring ----
  float* ring = a.ybuf + (uint64_t)stream * a.Rd;
  if (active) {
#pragma unroll
    for (int h = 0; h < RAD / 2; h++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        uint32_t idx = a.wpos + 2 * (tid + h * NT) + c;
        if (idx >= a.Rd) idx -= a.Rd;
        ring[idx] = o[2 * h + c];
      }
  }
  if constexpr (!FUSE_OUT) return;
  // ---- 6. PER_CHANNEL output stage: k_pcm_out's arithmetic for the one route of output `stream` ----
  __syncthreads();  // the block's ring writes are visible to the whole CTA
  if (!active) return;
  const RouteEntry en = a.entry[stream];
  const uint32_t obps = fmt_bytes(a.outfmt);
  const float inc = 1.0f / (float)M;
  // four samples at a time: their ring reads (and the 14-tap double-precision chains of the fractional mode) overlap
#pragma unroll 4
  for (int r = 0; r < RAD; r++) {
    const uint32_t n = tid + r * NT;
    float bus = 0.f;
    if (en.gain != 0.0f) {
      float v = delayed_read(ring, a.Rd, a.wpos, n, en.dcur, en.icur, a.fractional);
      if (en.flags & 1u) {
        const float vo = delayed_read(ring, a.Rd, a.wpos, n, en.dold, en.iold, a.fractional);
        const float g = __fmul_rn((float)n, inc);
        v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, g), vo), __fmul_rn(g, v));
      }
      bus = __fadd_rn(bus, __fmul_rn(en.gain, v));
    }
    store_from_f32(a.pcm_out + ((uint64_t)n * a.out_channels + stream) * obps, bus, a.outfmt, a.out_be != 0, a.out_fast != 0);
  }
}

}  // namespace bbx
