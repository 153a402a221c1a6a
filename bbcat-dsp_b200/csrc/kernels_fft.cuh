// kernels_fft.cuh -- forward transforms into the FDL ring (k_rfft*), partial sums -> inverse transforms -> crossfade -> delay
// ring (k_irfft*), and the spectrum exchange kernels of the input-sharded MIMO engine (gather for NCCL, peer-memory mixdown).
#pragma once

#include "kernels_common.cuh"
#include "async_copy.cuh"
#include "fft.cuh"

namespace bbx {

// ------------------------------------------------------------------------------------------
// k_rfft : windows of 2B floats -> packed spectra
// ------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB)
k_rfft(const float* __restrict__ src, uint64_t ch_stride, uint32_t win_stride, float2* __restrict__ dst, uint64_t dst_ch_stride,
       uint32_t R, uint32_t slot0, const float2* __restrict__ tw, float scale, uint32_t nch) {
  constexpr int RAD = FftCfg<M>::R, NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP;
  __shared__ float2 smem[FPB][MP];
  float2* s = smem[threadIdx.y];
  const int tid = threadIdx.x;
  const uint32_t chq = blockIdx.x * FPB + threadIdx.y, t = blockIdx.y;
  const bool active = chq < nch;
  const uint32_t ch = active ? chq : nch - 1;  // idle transforms of the last CTA recompute a valid one, stores masked
  const float2* win = reinterpret_cast<const float2*>(src + ch * ch_stride + (uint64_t)t * win_stride);
#pragma unroll
  for (int h = 0; h < RAD; h++) s[PADM<M>(tid + h * NT)] = win[tid + h * NT];
  __syncthreads();
  cfft_smem<M, false>(s, tw, tid);
  const uint32_t slot = (slot0 + t) % R;
  rfft_split_store<M>(s, tw, dst + ch * dst_ch_stride + (uint64_t)slot * M, scale, tid, active);
}

// Radix-8 sizes: persistent CTAs loop over (channel group, block) items with their twiddles in registers, the next
// item's window prefetched into registers, the first pass straight from those registers, and (for M = 512) one named
// barrier per transform instead of the block barrier.
template <int M>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB, (FftCfg<M>::NT * FftCfg<M>::FPB <= 256) ? 2 : 1)
k_rfft8(const float* __restrict__ src, uint64_t ch_stride, uint32_t win_stride, float2* __restrict__ dst, uint64_t dst_ch_stride,
        uint32_t R, uint32_t slot0, const float2* __restrict__ tw, float scale, uint32_t nch, uint32_t T) {
  constexpr int NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP;
  __shared__ float2 smem[FPB][MP];
  float2* s = smem[threadIdx.y];
  const int tid = threadIdx.x;
  Tw8<M> tw8;
  load_tw8<M>(tw8, tw, tid);
  const uint32_t nchg = ceil_div_dev(nch, (uint32_t)FPB), nitems = nchg * T;
  auto window = [&](uint32_t item, uint32_t& ch, uint32_t& t, bool& active) {
    t = item / nchg;
    const uint32_t chq = (item - t * nchg) * FPB + threadIdx.y;
    active = chq < nch;
    ch = active ? chq : nch - 1;  // idle transforms of a last group recompute a valid one, stores masked
    return reinterpret_cast<const float2*>(src + ch * ch_stride + (uint64_t)t * win_stride);
  };
  uint32_t item = blockIdx.x, ch = 0, t = 0;
  bool active = false;
  float2 vn[8];
  if (item < nitems) {
    const float2* win = window(item, ch, t, active);
#pragma unroll
    for (int r = 0; r < 8; r++) vn[r] = win[tid + r * NT];
  }
  for (; item < nitems; item += gridDim.x) {
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = vn[r];
    const uint32_t ch_c = ch, t_c = t;
    const bool active_c = active;
    if (item + gridDim.x < nitems) {
      const float2* win = window(item + gridDim.x, ch, t, active);
#pragma unroll
      for (int r = 0; r < 8; r++) vn[r] = win[tid + r * NT];
    }
    fft_bar<M>();  // the previous item's split stage is done with the workspace
    pass8_first<M, false>(v, s, tid);
    passes8_rest<M, false>(s, tw8, tid);
    const uint32_t slot = (slot0 + t_c) % R;
    rfft_split_store8<M>(s, tw8, dst + ch_c * dst_ch_stride + (uint64_t)slot * M, scale, tid, active_c);
  }
}

// ------------------------------------------------------------------------------------------
// k_irfft : partial sums -> time domain -> crossfade -> delay ring
// ------------------------------------------------------------------------------------------
struct PlanView {
  const uint32_t* job_slot_first;
  const uint32_t* job_slot_count;
  const uint32_t* xjob;  // per stream: extra job to crossfade into, kNoJob, or kSameJob
};

template <int M>
__device__ __forceinline__ void job_to_block(const float2* __restrict__ ypart_t, const float* __restrict__ nyq_t, uint64_t s_stride,
                                             uint32_t first, uint32_t count, float2* __restrict__ x,
                                             float2* __restrict__ s, const float2* __restrict__ tw, int tid,
                                             float (&o)[FftCfg<M>::R]) {
  constexpr int RAD = FftCfg<M>::R, NT = FftCfg<M>::NT;
#pragma unroll
  for (int h = 0; h < RAD; h++) {
    const int k = tid + h * NT;
    float2 a = make_float2(0.f, 0.f);
    // fixed slot order (deterministic sums); four loads in flight per step
    for (uint32_t sl = 0; sl < count; sl += 4) {
      float2 v[4];
#pragma unroll
      for (int q = 0; q < 4; q++)
        v[q] = (sl + q < count) ? ypart_t[(uint64_t)(first + sl + q) * s_stride + k] : make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (sl + q < count) {
          a.x += v[q].x;
          a.y += v[q].y;
        }
    }
    if (k == 0 && nyq_t) {
      // bin 0: the MAC kernels left G = DC - N in the real part; add the Nyquist sum back (same slot order)
      // (nyq_t == NULL: the tensor-core MIMO path writes bin 0 = (DC, Nyquist) directly)
      float n = 0.f;
      for (uint32_t sl = 0; sl < count; sl++) n += nyq_t[first + sl];
      a = make_float2(a.x + n, n);
    }
    x[k] = a;
  }
  __syncthreads();
  irfft_unsplit<M>(x, tw, s, tid);
  __syncthreads();
  cfft_smem<M, true>(s, tw, tid);
  // overlap-save: y[B+n] = component (n&1) of z[M/2 + n/2]; this thread keeps n = 2(tid + h NT) + {0,1}, h < R/2
#pragma unroll
  for (int h = 0; h < RAD / 2; h++) {
    float2 z = s[PADM<M>(M / 2 + tid + h * NT)];
    o[2 * h] = z.x;
    o[2 * h + 1] = z.y;
  }
  __syncthreads();
}

template <int M>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB)
k_irfft(const float2* __restrict__ ypart, uint32_t slot_stride, PlanView first_blk, PlanView steady, uint32_t n_first,
        const float2* __restrict__ tw, float* __restrict__ ybuf, uint32_t Rd, uint32_t wpos0, uint32_t n_streams,
        const float* __restrict__ nyq_part, uint64_t t_stride, uint64_t s_stride, uint32_t stream0) {
  // spectrum of (block t, slot s) at ypart + t * t_stride + s * s_stride (the MAC kernels: t_stride = slot_stride * M,
  // s_stride = M; after the sharded reduce-scatter: [slot][t][M]); streams stream0 .. stream0 + n_streams - 1
  constexpr int RAD = FftCfg<M>::R, NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP;
  extern __shared__ float2 k_irfft_smem[];  // per transform: summed spectrum x[M] + padded FFT workspace s[MP]
  float2* x = k_irfft_smem + (size_t)threadIdx.y * (M + MP);
  float2* s = x + M;
  const int tid = threadIdx.x;
  const uint32_t sq = blockIdx.x * FPB + threadIdx.y, t = blockIdx.y;
  const bool active = sq < n_streams;
  const uint32_t stream = stream0 + (active ? sq : n_streams - 1);
  const bool first = t < n_first;  // blocks covered by the transitional plan (0 or 1 of them)
  const PlanView pv = first ? first_blk : steady;
  const float2* ypart_t = ypart + (uint64_t)t * t_stride;
  const float* nyq_t = nyq_part ? nyq_part + (uint64_t)t * slot_stride : nullptr;
  float o[RAD];
  job_to_block<M>(ypart_t, nyq_t, s_stride, pv.job_slot_first[stream], pv.job_slot_count[stream], x, s, tw, tid, o);
  // the crossfade decision must be uniform across the CTA (block-wide barriers inside job_to_block):
  // every transform of the CTA runs the second pass when any of them needs it
  const uint32_t xj = first ? pv.xjob[stream] : kNoJob;
  const int any_x = __syncthreads_or(xj != kNoJob && xj != kSameJob);
  float o2[RAD];
  if (any_x) {
    const bool mine = (xj != kNoJob && xj != kSameJob);
    job_to_block<M>(ypart_t, nyq_t, s_stride, mine ? pv.job_slot_first[xj] : 0u, mine ? pv.job_slot_count[xj] : 0u, x, s, tw, tid, o2);
  }
  if (xj != kNoJob) {
    if (xj == kSameJob) {
#pragma unroll
      for (int i = 0; i < RAD; i++) o2[i] = o[i];
    }
    // out = (1-g) o_f + g o_f', g_n = n/B  (MixSamples + Interpolator ramp, sampled before the step)
    const float inc = 1.0f / (float)M;
#pragma unroll
    for (int h = 0; h < RAD / 2; h++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const uint32_t n = 2 * (tid + h * NT) + c;
        const float g = __fmul_rn((float)n, inc);
        const float a = __fmul_rn(__fsub_rn(1.0f, g), o[2 * h + c]);
        const float b = __fmul_rn(g, o2[2 * h + c]);
        o[2 * h + c] = __fadd_rn(a, b);
      }
  }
  if (!active) return;
  float* ring = ybuf + (uint64_t)stream * Rd;
  const uint32_t w = (wpos0 + t * (uint32_t)M) % Rd;
#pragma unroll
  for (int h = 0; h < RAD / 2; h++)
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const uint32_t n = 2 * (tid + h * NT) + c;
      uint32_t idx = w + n;
      if (idx >= Rd) idx -= Rd;
      ring[idx] = o[2 * h + c];
    }
}

// Radix-8 sizes: same contract as job_to_block, with cached twiddles, all of a slot's partial loads in flight at once,
// the inverse split and the first pass in registers, per-transform barriers.
template <int M>
__device__ __forceinline__ void job_to_block8(const float2* __restrict__ ypart_t, const float* __restrict__ nyq_t, uint64_t s_stride,
                                              uint32_t first, uint32_t count, float2* __restrict__ x, float2* __restrict__ s,
                                              const Tw8<M>& tw8, int tid, float (&o)[8]) {
  constexpr int NT = M / 8;
  float2 a[8];
#pragma unroll
  for (int h = 0; h < 8; h++) a[h] = make_float2(0.f, 0.f);
  for (uint32_t sl = 0; sl < count; sl++) {  // fixed slot order (deterministic sums)
    const float2* row = ypart_t + (uint64_t)(first + sl) * s_stride;
    float2 v[8];
#pragma unroll
    for (int h = 0; h < 8; h++) v[h] = row[tid + h * NT];
#pragma unroll
    for (int h = 0; h < 8; h++) {
      a[h].x += v[h].x;
      a[h].y += v[h].y;
    }
  }
  if (tid == 0 && nyq_t) {
    // bin 0: the MAC kernels left G = DC - N in the real part; add the Nyquist sum back (same slot order)
    float n = 0.f;
    for (uint32_t sl = 0; sl < count; sl++) n += nyq_t[first + sl];
    a[0] = make_float2(a[0].x + n, n);
  }
#pragma unroll
  for (int h = 0; h < 8; h++) x[tid + h * NT] = a[h];
  fft_bar<M>();  // x complete; also: every thread of the transform is past its reads of s from the previous job
  float2 v[8];
  irfft_unsplit8<M>(x, tw8, v, tid);
  pass8_first<M, true>(v, s, tid);
  passes8_rest<M, true>(s, tw8, tid);
  // overlap-save: y[B+n] = component (n&1) of z[M/2 + n/2]; this thread keeps n = 2(tid + h NT) + {0,1}, h < 4
#pragma unroll
  for (int h = 0; h < 4; h++) {
    float2 z = s[PADM<M>(M / 2 + tid + h * NT)];
    o[2 * h] = z.x;
    o[2 * h + 1] = z.y;
  }
}

// The persistent inverse kernel keeps the partial-sum rows of its NEXT item in flight while it transforms the current one:
// every thread copies exactly the elements it will add (8-byte cp.async into a per-transform staging area, up to
// kIrfftPre slots of a job; further slots are read directly), so no barrier stands between the copies and their use, and
// the staging area is free again as soon as the thread has the sums in registers.  Without it the kernel alternated
// between a load phase with 64 bytes in flight per thread and a transform phase with none (34 us per 64-block C3 step for
// 90 MB of traffic).
static constexpr int kIrfftPre = 3;
template <int M>
constexpr size_t irfft8_smem_bytes() {
  return sizeof(float2) * (size_t)FftCfg<M>::FPB * (M + FftCfg<M>::MP + kIrfftPre * M) + sizeof(float) * FftCfg<M>::FPB * 4;
}

template <int M>
__global__ void __launch_bounds__(FftCfg<M>::NT * FftCfg<M>::FPB, (FftCfg<M>::NT * FftCfg<M>::FPB <= 256) ? 2 : 1)
k_irfft8(const float2* __restrict__ ypart, uint32_t slot_stride, PlanView first_blk, PlanView steady, uint32_t n_first,
         const float2* __restrict__ tw, float* __restrict__ ybuf, uint32_t Rd, uint32_t wpos0, uint32_t n_streams,
         const float* __restrict__ nyq_part, uint64_t t_stride, uint64_t s_stride, uint32_t stream0, uint32_t T) {
  constexpr int RAD = 8, NT = FftCfg<M>::NT, FPB = FftCfg<M>::FPB, MP = FftCfg<M>::MP, PRE = kIrfftPre;
  constexpr int PER = M + MP + PRE * M;
  extern __shared__ float2 k_irfft_smem[];  // per transform: summed spectrum x[M], padded FFT workspace s[MP], staging pre[PRE][M]
  float2* x = k_irfft_smem + (size_t)threadIdx.y * PER;
  float2* s = x + M;
  const float2* pre = s + MP;
  const float* pre_nyq = reinterpret_cast<const float*>(k_irfft_smem + (size_t)FPB * PER) + threadIdx.y * 4;
  const uint32_t pre_sm = (uint32_t)__cvta_generic_to_shared(pre), nyq_sm = (uint32_t)__cvta_generic_to_shared(pre_nyq);
  const int tid = threadIdx.x;
  Tw8<M> tw8;
  load_tw8<M>(tw8, tw, tid);
  const uint32_t nsg = ceil_div_dev(n_streams, (uint32_t)FPB), nitems = nsg * T;
  struct Item {
    uint32_t t, stream, first, count;
    bool active, firstblk;
  };
  auto describe = [&](uint32_t item) {
    Item it;
    it.t = item / nsg;
    const uint32_t sq = (item - it.t * nsg) * FPB + threadIdx.y;
    it.active = sq < n_streams;
    it.stream = stream0 + (it.active ? sq : n_streams - 1);
    it.firstblk = it.t < n_first;  // blocks covered by the transitional plan (0 or 1 of them)
    const PlanView& pv = it.firstblk ? first_blk : steady;
    it.first = pv.job_slot_first[it.stream];
    it.count = pv.job_slot_count[it.stream];
    return it;
  };
  auto prefetch = [&](const Item& it) {
    const uint32_t npre = min(it.count, (uint32_t)PRE);
    const float2* row = ypart + (uint64_t)it.t * t_stride + (uint64_t)it.first * s_stride + tid;
    for (uint32_t sl = 0; sl < npre; sl++, row += s_stride) {
#pragma unroll
      for (int h = 0; h < 8; h++) ac::cp_async8(pre_sm + (sl * M + tid + h * NT) * 8u, row + h * NT);
    }
    if (tid == 0 && nyq_part)
      for (uint32_t sl = 0; sl < npre; sl++) ac::cp_async4(nyq_sm + sl * 4u, nyq_part + (uint64_t)it.t * slot_stride + it.first + sl);
    ac::cp_async_commit();
  };
  uint32_t item = blockIdx.x;
  Item cur = Item();
  if (item < nitems) {
    cur = describe(item);
    prefetch(cur);
  }
  for (; item < nitems; item += gridDim.x) {
    const bool more = item + gridDim.x < nitems;
    Item nxt = cur;
    if (more) nxt = describe(item + gridDim.x);  // plan entries of the next item: loaded under the wait below
    const float2* ypart_t = ypart + (uint64_t)cur.t * t_stride;
    const float* nyq_t = nyq_part ? nyq_part + (uint64_t)cur.t * slot_stride : nullptr;
    // ---- sums of the job's slots in fixed order: staged slots first, the rest (jobs of more than PRE slots) directly ----
    float2 a[8];
#pragma unroll
    for (int h = 0; h < 8; h++) a[h] = make_float2(0.f, 0.f);
    float n = 0.f;
    ac::cp_async_wait_all();
    const uint32_t npre = min(cur.count, (uint32_t)PRE);
    for (uint32_t sl = 0; sl < npre; sl++) {
      float2 v[8];
#pragma unroll
      for (int h = 0; h < 8; h++) v[h] = pre[sl * M + tid + h * NT];
#pragma unroll
      for (int h = 0; h < 8; h++) {
        a[h].x += v[h].x;
        a[h].y += v[h].y;
      }
      if (tid == 0 && nyq_t) n += pre_nyq[sl];
    }
    for (uint32_t sl = npre; sl < cur.count; sl++) {
      const float2* row = ypart_t + (uint64_t)(cur.first + sl) * s_stride;
      float2 v[8];
#pragma unroll
      for (int h = 0; h < 8; h++) v[h] = row[tid + h * NT];
#pragma unroll
      for (int h = 0; h < 8; h++) {
        a[h].x += v[h].x;
        a[h].y += v[h].y;
      }
      if (tid == 0 && nyq_t) n += nyq_t[cur.first + sl];
    }
    // bin 0: the MAC kernels left G = DC - N in the real part; add the Nyquist sum back (same slot order)
    if (tid == 0 && nyq_t) a[0] = make_float2(a[0].x + n, n);
    if (more) prefetch(nxt);  // this thread's staging elements are in registers now: the next item's rows go in flight
    // ---- summed spectrum -> time domain ----
    float o[RAD];
#pragma unroll
    for (int h = 0; h < 8; h++) x[tid + h * NT] = a[h];
    fft_bar<M>();  // x complete; also: every thread of the transform is past its reads of s from the previous item
    {
      float2 v[8];
      irfft_unsplit8<M>(x, tw8, v, tid);
      pass8_first<M, true>(v, s, tid);
      passes8_rest<M, true>(s, tw8, tid);
#pragma unroll
      for (int h = 0; h < 4; h++) {
        float2 z = s[PADM<M>(M / 2 + tid + h * NT)];
        o[2 * h] = z.x;
        o[2 * h + 1] = z.y;
      }
    }
    const PlanView& pv = cur.firstblk ? first_blk : steady;
    const uint32_t xj = cur.firstblk ? pv.xjob[cur.stream] : kNoJob;
    // the barriers inside job_to_block8 span one transform (M = 512) or the CTA: the decision to run the second
    // pass is made CTA-uniform so that both cases are safe
    const int any_x = n_first ? __syncthreads_or(xj != kNoJob && xj != kSameJob) : 0;
    float o2[RAD];
    if (any_x) {
      const bool mine = (xj != kNoJob && xj != kSameJob);
      job_to_block8<M>(ypart_t, nyq_t, s_stride, mine ? pv.job_slot_first[xj] : 0u, mine ? pv.job_slot_count[xj] : 0u, x, s, tw8,
                       tid, o2);
    }
    if (xj != kNoJob) {
      if (xj == kSameJob) {
#pragma unroll
        for (int i = 0; i < RAD; i++) o2[i] = o[i];
      }
      // out = (1-g) o_f + g o_f', g_n = n/B  (MixSamples + Interpolator ramp, sampled before the step)
      const float inc = 1.0f / (float)M;
#pragma unroll
      for (int h = 0; h < RAD / 2; h++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const uint32_t nn = 2 * (tid + h * NT) + c;
          const float g = __fmul_rn((float)nn, inc);
          const float aa = __fmul_rn(__fsub_rn(1.0f, g), o[2 * h + c]);
          const float bb = __fmul_rn(g, o2[2 * h + c]);
          o[2 * h + c] = __fadd_rn(aa, bb);
        }
    }
    if (cur.active) {
      float* ring = ybuf + (uint64_t)cur.stream * Rd;
      const uint32_t w = (wpos0 + cur.t * (uint32_t)M) % Rd;
#pragma unroll
      for (int h = 0; h < RAD / 2; h++) {
        const uint32_t nn = 2 * (tid + h * NT);
        uint32_t idx = w + nn;  // w and nn are even, Rd is a multiple of the block size: the pair never straddles the wrap
        if (idx >= Rd) idx -= Rd;
        if ((Rd & 1u) == 0 && (w & 1u) == 0) {
          *reinterpret_cast<float2*>(ring + idx) = make_float2(o[2 * h], o[2 * h + 1]);
        } else {
          ring[idx] = o[2 * h];
          uint32_t i1 = idx + 1;
          if (i1 >= Rd) i1 -= Rd;
          ring[i1] = o[2 * h + 1];
        }
      }
    }
    cur = nxt;
  }
}

// Input-sharded MIMO: this rank's partial output spectra, job by job (slots summed in fixed order, bin 0 restored to
// (DC, Nyquist)), into the reduce-scatter send buffer [output][t][B] -- one contiguous chunk per destination rank.
__global__ void __launch_bounds__(256) k_gather_spectra(const float2* __restrict__ ypart, const float* __restrict__ nyq_part,
                                                        uint32_t slot_stride, PlanView pv, float2* __restrict__ send, uint32_t B,
                                                        uint32_t T) {
  const uint32_t o = blockIdx.x, t = blockIdx.y;
  const uint32_t first = pv.job_slot_first[o], count = pv.job_slot_count[o];
  const float2* yt = ypart + (uint64_t)t * slot_stride * B;
  for (uint32_t k = threadIdx.x; k < B; k += blockDim.x) {
    float2 a = make_float2(0.f, 0.f);
    for (uint32_t sl = 0; sl < count; sl++) {
      const float2 v = yt[(uint64_t)(first + sl) * B + k];
      a.x += v.x;
      a.y += v.y;
    }
    if (k == 0 && nyq_part) {
      float n = 0.f;
      for (uint32_t sl = 0; sl < count; sl++) n += nyq_part[(uint64_t)t * slot_stride + first + sl];
      a = make_float2(a.x + n, n);
    }
    send[((uint64_t)o * T + t) * B + k] = a;
  }
}

// ---- peer-memory mixdown: the reduce of the input-sharded MIMO engine without a collective library ----------------------
// Every rank adds its partial slots like k_gather_spectra, but stores the spectrum of output o straight into the memory of
// the rank that owns o (NVLink peer stores into a buffer opened with cudaIpcOpenMemHandle), at slot (o_local, source rank).
// The owner's inverse-transform kernel then adds the `world` slots of an output in rank order -- the same fixed-order slot
// sum it already runs over the MAC's partial sums, so the result does not depend on a collective's reduction schedule.
// Completion: the last CTA of a launch publishes the call's epoch in every peer's flag array after a system-scope fence;
// k_peer_wait (one warp, in stream order before the inverse transforms) spins until all sources have published it.  The
// receive buffer is double-buffered by epoch parity: a source can only be two calls ahead after it has seen this rank's
// flag of the call in between, which this rank publishes after its own inverse transforms of the older call (stream order).
struct PeerTable {
  float2* data[16];     // receive buffers of the ranks (own rank: the local buffer)
  uint32_t* flags[16];  // their flag arrays, [2][world]
};

// One CTA = one output x kPeerTPB block-steps: 16-byte peer stores, and one system-scope fence + one atomic per CTA for eight
// rows instead of one (round 1 launched a CTA per row: 4096 fences and atomics per exchange, 62 us for 14.7 MB at 8 ranks).
static constexpr uint32_t kPeerTPB = 8;
__global__ void __launch_bounds__(256) k_gather_spectra_peer(const float2* __restrict__ ypart, const float* __restrict__ nyq_part,
                                                             uint32_t slot_stride, PlanView pv, PeerTable pt, uint32_t world,
                                                             uint32_t rank, uint32_t nloc, uint32_t B, uint32_t T, uint64_t half,
                                                             uint32_t parity, uint32_t epoch, uint32_t* __restrict__ done) {
  const uint32_t o = blockIdx.x, t0 = blockIdx.y * kPeerTPB, nt = min(kPeerTPB, T - t0);
  const uint32_t first = pv.job_slot_first[o], count = pv.job_slot_count[o];
  const uint32_t r = o / nloc, ol = o - r * nloc;
  float2* dst0 = pt.data[r] + (uint64_t)parity * half + (((uint64_t)ol * world + rank) * T + t0) * B;
  const uint32_t halfB = B / 2;  // float4 = two bins
  for (uint32_t idx = threadIdx.x; idx < nt * halfB; idx += blockDim.x) {
    const uint32_t tt = idx / halfB, k4 = idx - tt * halfB;
    const float4* yt = reinterpret_cast<const float4*>(ypart + (uint64_t)(t0 + tt) * slot_stride * B);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t sl = 0; sl < count; sl++) {
      const float4 v = yt[(uint64_t)(first + sl) * halfB + k4];
      a.x += v.x;
      a.y += v.y;
      a.z += v.z;
      a.w += v.w;
    }
    if (k4 == 0 && nyq_part) {
      float n = 0.f;
      for (uint32_t sl = 0; sl < count; sl++) n += nyq_part[(uint64_t)(t0 + tt) * slot_stride + first + sl];
      a.x += n;
      a.y = n;
    }
    reinterpret_cast<float4*>(dst0 + (uint64_t)tt * B)[k4] = a;
  }
  __threadfence_system();  // this thread's peer stores are performed before the CTA counts itself done
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t total = gridDim.x * gridDim.y;
    if (atomicAdd(done, 1u) == total - 1) {
      *done = 0;  // ready for the next launch (stream-ordered)
      __threadfence_system();
      for (uint32_t q = 0; q < world; q++) *((volatile uint32_t*)pt.flags[q] + parity * world + rank) = epoch;
    }
  }
}

__global__ void __launch_bounds__(32) k_peer_wait(const uint32_t* flags, uint32_t world, uint32_t parity, uint32_t epoch,
                                                  int* status) {
  if (threadIdx.x < world) {
    const volatile uint32_t* f = flags + parity * world + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(*f - epoch) < 0) {
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a source never arrived; bbx_engine_sync reports it
        *status = 1 + (int)threadIdx.x;
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

// comm.cu

}  // namespace bbx
