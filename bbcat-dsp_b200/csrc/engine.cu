// engine.cu -- the partitioned-convolution engine of libbbx (BlockConvolver / Convolver path).
//
// The reference's BlockConvolver.{h,cpp}, Convolver.{h,cpp}, FFT*.cpp and simd_utils (README:38-51,
// 68-69) are absent from the mounted tree; behaviour follows SURVEY.md 8.A.  Data flow of one
// bbx_process call over T blocks of B frames (all kernels on the engine stream; the kernels live in kernels_pcm.cuh,
// kernels_fft.cuh, kernels_mac.cuh and mimo_tc.cuh, this file is the host side: plans, routes, launches, the C ABI):
//
//   k_pcm_in   interleaved PCM (any SampleFormat_t, LE/BE) -> planar fp32 xin[input][(T+1)B]
//              (block 0 of the row is the previous call's last block = the overlap-save history)
//   k_rfft     window [prev | cur] (2B floats, contiguous in xin) -> packed spectrum -> FDL ring slot
//              FDL[input][(head+t) mod R][B] complex, R >= Pmax + Tmax - 1
//   k_fdl_mac  Y[job] = sum over (term, p) H[term][p] * FDL[input(term)][(head+t-p) mod R]
//              the hot kernel: streams H and FDL rows (8B bytes each, 4 KB at B=512) with 128-bit
//              loads, flattened (job, term, p) row space split evenly over a grid sized to the
//              SM count, per-CTA partial sums written to Ypart (deterministic, no atomics)
//   k_irfft    sum the partials in fixed order -> C2R -> keep samples B..2B-1 -> filter crossfade
//              -> per-stream delay ring ybuf[stream][Rd]
//   k_pcm_out  delayed ring reads (integer or 14-tap fractional, delay crossfade) -> MixSamples-order
//              mixdown -> float -> output SampleFormat_t, interleaved
//
// HBM layout: H per filter [P][B] float2 (1/N folded in), FDL [n_in][R][B] float2, both rows
// 16-byte aligned so one thread owns one float4 column (2 bins) across all partitions.
#include <math.h>

#include <algorithm>
#include <mutex>
#include <unordered_set>
#include <vector>

#include <time.h>

#include <nvtx3/nvToolsExt.h>  // header-only; a no-op (one pointer test per call) unless a profiler is attached

#include "common.cuh"
#include "kernels_common.cuh"
#include "kernels_fft.cuh"
#include "kernels_mac.cuh"
#include "kernels_pcm.cuh"
#include "kernels_fused.cuh"
#include "mac_tbs.h"
#include "mimo_tc.cuh"

namespace bbx {

__global__ void k_flush(float4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

int comm_reduce_scatter_f32(bbx_comm* c, const float* send, float* recv, size_t recvcount, cudaStream_t st);
int comm_world(const bbx_comm* c);
int comm_rank(const bbx_comm* c);

}  // namespace bbx

using namespace bbx;

// ==========================================================================================
// host side
// ==========================================================================================
// NVTX range around a stage of a call (host side: the range covers the stage's launches, the tool correlates the kernels)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct bbx_filter {
  bbx_engine* engine;
  float2* H;  // device [P][B]
  uint32_t P;
};

struct PathState {
  uint32_t input = 0, output = 0;
  float gain = 1.0f;
  const bbx_filter* cur = nullptr;
  const bbx_filter* pend = nullptr;
  bool has_pending = false, xfade = false;
  double delay = 0.0, pend_delay = 0.0;
};

// one job = one accumulated output spectrum: a list of (filter, input) terms
struct JobTerm {
  const bbx_filter* f;
  uint32_t input;
};

// Ring of pinned upload buffers for one device table.  Every upload is a copy on the engine stream followed by an event;
// re-using a slot waits only for the copy that last read it (kDepth uploads ago), so the host can prepare the plans and
// routes of the next calls while the device still works on the previous ones (dynamic IR switching: one upload set per
// switch).
struct Staging {
  static constexpr int kDepth = 4;
  uint8_t* h[kDepth] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[kDepth] = {nullptr, nullptr, nullptr, nullptr};
  bool pending[kDepth] = {false, false, false, false};
  int cur = 0;
  size_t bytes = 0;
};

// device-resident MAC plan + host mirror
struct MacPlan {
  // host staging (pinned ring) and device blob, same layout
  Staging stg;
  uint8_t* h_blob = nullptr;  // the slot being filled by build_plan
  uint8_t* d_blob = nullptr;
  size_t blob_bytes = 0;
  // offsets inside the blob
  size_t off_segs = 0, off_cta = 0, off_first = 0, off_count = 0, off_xjob = 0, off_jseg = 0;
  uint32_t n_ctas = 0, n_slots = 0, n_jobs = 0, total_rows = 0, n_terms = 0;
  uint32_t max_job_rows = 0;  // longest job in partitions (the fused single-launch path walks a job inside one CTA)
  uint32_t occ = 1;  // streaming-MAC variant this plan was cut for (CTAs per SM <-> unroll depth)
  bool valid = false;
  const MacSeg* segs() const { return (const MacSeg*)(d_blob + off_segs); }
  const uint32_t* cta_seg_begin() const { return (const uint32_t*)(d_blob + off_cta); }
  const uint32_t* job_seg_first() const { return (const uint32_t*)(d_blob + off_jseg); }
  PlanView view() const {
    PlanView v;
    v.job_slot_first = (const uint32_t*)(d_blob + off_first);
    v.job_slot_count = (const uint32_t*)(d_blob + off_count);
    v.xjob = (const uint32_t*)(d_blob + off_xjob);
    return v;
  }
};

struct bbx_engine {
  bbx_config cfg;
  int device = 0;
  uint32_t B = 0, Pmax = 0, n_in = 0, n_out = 0, n_paths = 0, n_streams = 0, Tmax = 1;
  uint32_t R = 0;   // FDL ring slots
  uint32_t Rd = 0;  // delay ring frames
  uint32_t xstride = 0;
  int mode = 0;
  cudaStream_t stream = nullptr;
  // device buffers
  float2* tw = nullptr;
  float* xin[2] = {nullptr, nullptr};
  float2* fdl = nullptr;
  float2* ypart = nullptr;
  float* nyq_part = nullptr;  // [Tmax][max_slots] Nyquist partial sums (see cmac)
  float* ybuf = nullptr;
  // host-pointer path: PCM staging in kIoSlots buffers per direction, copies on their own streams so that the H2D of
  // call n+1 and the D2H of call n-1 overlap the kernels of call n.  Three slots, not two: with two, every call sits on
  // the cycle "kernels of n-2 done -> H2D of n -> kernels of n" (and the same through the D2H side), so each
  // cross-stream hand-off (5..30 us on this part, tools/e2e_timeline.py) adds to the step; with three the copies run
  // back to back and the step is the copy time.
  static constexpr int kIoSlots = 3;
  uint8_t* d_in[kIoSlots] = {};
  uint8_t* d_out[kIoSlots] = {};
  size_t d_io_bytes = 0;
  uint8_t *d_lat_in = nullptr, *d_lat_out = nullptr;  // staging of small (latency-mode) host calls, used on the engine stream
  size_t d_lat_bytes = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaStream_t s_aux = nullptr;  // side stream: k_nyq_mac runs next to the time-batched MAC
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_h2d[kIoSlots] = {}, ev_comp[kIoSlots] = {}, ev_d2h[kIoSlots] = {};
  cudaEvent_t ev_join_in = nullptr, ev_join_out = nullptr;
  uint64_t host_calls = 0;
  // optional trace of the host-buffer pipeline (bbx_engine_io_trace): six timing events per call
  std::vector<cudaEvent_t> io_ev;
  size_t io_calls = 0, io_cap = 0;
  // short host calls: the PCM kernels read / write the caller's pinned (device-mapped) buffers directly over PCIe
  // instead of going through the copy engines and the staging buffers; 0 disables
  size_t direct_io_max_bytes = 1u << 20;
  uint64_t direct_calls = 0;
  bool copy_streams_busy = true;  // work has been enqueued on s_in / s_out since the last full synchronise
  float4* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // route tables (device blob + pinned staging)
  Staging route_stg;
  uint32_t n_routes_pcm = 0;  // routes feeding the PCM outputs (set by upload_routes)
  bool pcm_out_mix = true;  // bbx_engine_set_mixdown_kernel: false keeps mixdowns on the per-output kernel
  uint8_t* h_route = nullptr;  // the slot being filled by upload_routes
  uint8_t* d_route = nullptr;
  size_t route_bytes = 0, roff_first = 0, roff_stream = 0, roff_gain = 0, roff_dcur = 0, roff_dold = 0, roff_flags = 0,
         roff_icur = 0, roff_iold = 0, roff_entry = 0, roff_input = 0;
  bool in_is_host = false, out_is_host = false;  // this call's PCM buffers are pinned host memory read / written in place
  bool fused_on = true;         // streaming calls of PER_CHANNEL / ROUTED engines with short filters: one launch (k_block_fused)
  uint32_t fused_max_rows = 32;  // ... "short" = at most this many partitions per path
  uint64_t fused_calls = 0;
  bool route_dirty = true;
  // plans
  MacPlan plan_first, plan_steady;
  uint32_t max_segs = 0, max_slots = 0, max_ctas = 0, max_jobs = 0;
  bool steady_dirty = true;
  // state
  std::vector<bbx_filter*> filters;  // every live filter of this engine (they die with it at the latest)
  std::vector<PathState> paths;
  uint32_t head = 0, wpos = 0, parity = 0, tprev = 1;
  // measurement
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  cudaEvent_t ev_upload = nullptr;  // last H2D copy out of the pinned plan/route staging
  bool upload_pending = false;
  uint32_t mac_occ = 1;          // resident streaming-MAC CTAs per SM; the plan has 148 * mac_occ row ranges
  float mac_l2_keep = 3.f / 16;  // fraction of H / FDL lines given L2 evict-last priority by the streaming MAC
  uint32_t mac_time_tile = 16;   // time-batched MAC: 16 = k_fdl_mac_tbs (shared operand stream), 116 / 32 = round 1's
                                 // k_fdl_mac_tb<16> / <32> (kept for A/B runs), 0 = streaming kernel only
  int* mac_status_h = nullptr;   // mapped pinned word k_fdl_mac_tbs sets when a barrier wait timed out
  int* mac_status = nullptr;
  uint64_t launches = 0;
  bool profile_mac = false;
  std::vector<cudaEvent_t> mac_events;  // pairs
  size_t mac_events_used = 0;
  std::vector<cudaEvent_t> xchg_events;  // pairs around the input-sharded MIMO exchange (gather + peer stores / reduce-scatter)
  size_t xchg_events_used = 0;
  double xchg_ms_total = 0.0;
  uint64_t xchg_count = 0, xchg_bytes = 0;
  double mac_ms_total = 0.0;
  uint64_t mac_launches = 0, mac_units = 0, mac_bytes = 0;
  int last_infmt = FMT_F32, last_outfmt = FMT_F32;
  const char* last_mac_kernel = "none";  // name of the MAC kernel of the most recent call (bbx_engine_mac_kernel)
  // input-sharded MIMO (bbx_config::mimo_shard_*): partial spectra -> reduce-scatter -> local outputs only
  uint32_t sh_world = 1, sh_rank = 0, sh_o0 = 0, sh_nloc = 0;  // local outputs [sh_o0, sh_o0 + sh_nloc)
  uint32_t n_out_pcm = 0;                                      // channels written by bbx_process (= sh_nloc when sharded)
  bbx_comm* comm = nullptr;
  // peer-memory mixdown (bbx_engine_peer_export / _attach): receive buffer [2][sh_nloc][world][Tmax][B] + flags [2][world]
  bool px_on = false;
  uint8_t* px_mem = nullptr;       // one allocation: data, then the flags
  size_t px_half = 0;              // float2 elements of one parity half
  size_t px_flag_off = 0;          // byte offset of the flags
  void* px_peer[16] = {nullptr};   // opened peer allocations (own rank: nullptr)
  PeerTable px_table;
  uint32_t px_epoch = 0;
  uint32_t* px_done = nullptr;     // last-CTA counter of k_gather_spectra_peer
  uint32_t* px_view = nullptr;     // device: first[n_out] | count[n_out] | xjob[n_out] for the world slots per local output
  int* px_status_h = nullptr;      // mapped pinned word: k_peer_wait timed out
  int* px_status = nullptr;
  float2* sh_send = nullptr;  // [n_out][T][B]
  float2* sh_recv = nullptr;  // [sh_nloc][T][B]
  uint32_t* sh_view = nullptr;  // device: first[n_out] | count[n_out] | xjob[n_out], first[o] = o - sh_o0
  // MIMO on the tensor cores (mimo_tc.cuh): bin-major operand copies and a one-slot-per-output plan view
  bool tc_on = false;          // buffers exist (MIMO mode, max_blocks >= tc_min_blocks, not disabled)
  bool tc_dirty = true;        // Hpack must be rebuilt from the current filter matrix
  uint32_t tc_min_blocks = 16; // calls with fewer blocks use the streaming SIMT MAC
  uint32_t tc_P2 = 1, tc_P2log = 0, tc_G = 0, tc_nog = 0;
  uint64_t tc_xbin_max = 0;  // float2 elements per bin of tc_xb at N = 64
  float4* tc_hpack = nullptr;
  float2* tc_xb = nullptr;
  const float2** tc_ftab_h = nullptr;  // pinned staging [n_out][n_in]
  const float2** tc_ftab_d = nullptr;
  uint32_t* tc_fparts_h = nullptr;
  uint32_t* tc_fparts_d = nullptr;
  uint32_t* tc_view = nullptr;  // device: first[n_out] | count[n_out] | xjob[n_out]
  int* tc_status = nullptr;    // device view of tc_status_h
  int* tc_status_h = nullptr;  // mapped pinned word the kernel sets when a barrier wait times out
  unsigned long long* tc_trace = nullptr;  // optional per-CTA role cycle counters (bbx_engine_tensor_trace)
  uint32_t tc_trace_ctas = 0;
  uint64_t tc_launches = 0;
};

namespace {

// Live filter handles of the process.  An engine releases the filters it still owns when it is destroyed; a filter
// handle destroyed after its engine (C++ destruction order, Python garbage collection) is then simply unknown here
// instead of a dangling pointer.
std::mutex g_filter_mu;
std::unordered_set<const bbx_filter*> g_live_filters;

// persistent grid of the radix-8 kernels: enough CTAs to fill the machine, never more than there are items
template <typename K>
uint32_t persistent_grid(K kernel, int threads, size_t smem, uint32_t nitems) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return std::max(1u, std::min(nitems, (uint32_t)(kNumSMs * per_sm)));
}

template <int M>
void launch_rfft_t(const float* src, uint64_t ch_stride, uint32_t win_stride, float2* dst, uint64_t dst_ch_stride, uint32_t R,
                   uint32_t slot0, const float2* tw, float scale, uint32_t nch, uint32_t T, cudaStream_t st) {
  constexpr int FPB = FftCfg<M>::FPB;
  if constexpr (FftCfg<M>::R == 8) {
    static uint32_t per_sm_grid = 0;  // occupancy query once per size
    if (!per_sm_grid) per_sm_grid = persistent_grid(k_rfft8<M>, FftCfg<M>::NT * FPB, 0, 1u << 30);
    const uint32_t nitems = ceil_div(nch, FPB) * T;
    k_rfft8<M><<<std::min(nitems, per_sm_grid), dim3(FftCfg<M>::NT, FPB), 0, st>>>(src, ch_stride, win_stride, dst, dst_ch_stride, R,
                                                                                  slot0, tw, scale, nch, T);
    return;
  }
  k_rfft<M><<<dim3(ceil_div(nch, FPB), T), dim3(FftCfg<M>::NT, FPB), 0, st>>>(src, ch_stride, win_stride, dst, dst_ch_stride, R,
                                                                              slot0, tw, scale, nch);
}

int launch_rfft(uint32_t B, const float* src, uint64_t ch_stride, uint32_t win_stride, float2* dst, uint64_t dst_ch_stride,
                uint32_t R, uint32_t slot0, const float2* tw, float scale, uint32_t nch, uint32_t T, cudaStream_t st) {
  switch (B) {
    case 64: launch_rfft_t<64>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 128: launch_rfft_t<128>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 256: launch_rfft_t<256>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 512: launch_rfft_t<512>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 1024: launch_rfft_t<1024>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 2048: launch_rfft_t<2048>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    case 4096: launch_rfft_t<4096>(src, ch_stride, win_stride, dst, dst_ch_stride, R, slot0, tw, scale, nch, T, st); break;
    default: set_error("unsupported block size %u", B); return BBX_ERR_UNSUPPORTED;
  }
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

// view of the tensor-core MIMO result: output o = slot o, one slot per job, no crossfade
PlanView tc_plan_view(const bbx_engine* e) {
  PlanView v;
  v.job_slot_first = e->tc_view;
  v.job_slot_count = e->tc_view + e->n_out;
  v.xjob = e->tc_view + 2 * (size_t)e->n_out;
  return v;
}

PlanView peer_plan_view(const bbx_engine* e) {
  PlanView v;
  v.job_slot_first = e->px_view;
  v.job_slot_count = e->px_view + e->n_out;
  v.xjob = e->px_view + 2 * (size_t)e->n_out;
  return v;
}

PlanView shard_plan_view(const bbx_engine* e) {
  PlanView v;
  v.job_slot_first = e->sh_view;
  v.job_slot_count = e->sh_view + e->n_out;
  v.xjob = e->sh_view + 2 * (size_t)e->n_out;
  return v;
}

// persistent grid of k_irfft8<M> on the current device; its dynamic shared memory (workspace + staging of the next item) is
// above the 48 KB default, and function attributes belong to a device's context: set once per device, not once per process
template <int M>
uint32_t irfft8_grid() {
  static uint32_t grid[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!grid[dev]) {
    constexpr size_t smem8 = irfft8_smem_bytes<M>();
    if (smem8 > 48 * 1024) cudaFuncSetAttribute(k_irfft8<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8);
    grid[dev] = persistent_grid(k_irfft8<M>, FftCfg<M>::NT * FftCfg<M>::FPB, smem8, 1u << 30);
  }
  return grid[dev];
}

template <int M>
void launch_irfft_t(bbx_engine* e, uint32_t T, uint32_t n_first, bool tc, cudaStream_t st) {
  const float2* ypart = e->ypart;
  const float* nyq_part = e->nyq_part;
  const uint32_t wpos = e->wpos;
  constexpr int FPB = FftCfg<M>::FPB;
  constexpr size_t smem = sizeof(float2) * (size_t)FPB * (M + FftCfg<M>::MP);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_irfft<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e->sh_world > 1 || e->comm) {
    // reduced spectra of the local outputs, [local output][t][M]; peer mode: [local output][source rank][t][M], the
    // kernel adds the `world` slots of an output in rank order
    const PlanView v = e->px_on ? peer_plan_view(e) : shard_plan_view(e);
    const float2* spectra = e->px_on ? (const float2*)e->px_mem + (uint64_t)(e->px_epoch & 1u) * e->px_half : e->sh_recv;
    if constexpr (FftCfg<M>::R == 8) {
      constexpr size_t smem8 = irfft8_smem_bytes<M>();
      const uint32_t per_sm_grid = irfft8_grid<M>();
      const uint32_t nitems = ceil_div(e->sh_nloc, FPB) * T;
      k_irfft8<M><<<std::min(nitems, per_sm_grid), dim3(FftCfg<M>::NT, FPB), smem8, st>>>(
          spectra, e->max_slots, v, v, 0, e->tw, e->ybuf, e->Rd, e->wpos, e->sh_nloc, nullptr, (uint64_t)M, (uint64_t)T * M,
          e->sh_o0, T);
      return;
    }
    k_irfft<M><<<dim3(ceil_div(e->sh_nloc, FPB), T), dim3(FftCfg<M>::NT, FPB), smem, st>>>(
        spectra, e->max_slots, v, v, 0, e->tw, e->ybuf, e->Rd, e->wpos, e->sh_nloc, nullptr, (uint64_t)M, (uint64_t)T * M,
        e->sh_o0);
    return;
  }
  const PlanView first = tc ? tc_plan_view(e) : e->plan_first.view();
  const PlanView steady = tc ? tc_plan_view(e) : e->plan_steady.view();
  if constexpr (FftCfg<M>::R == 8) {
    constexpr size_t smem8 = irfft8_smem_bytes<M>();
    const uint32_t per_sm_grid = irfft8_grid<M>();
    const uint32_t nitems = ceil_div(e->n_streams, FPB) * T;
    k_irfft8<M><<<std::min(nitems, per_sm_grid), dim3(FftCfg<M>::NT, FPB), smem8, st>>>(
        ypart, e->max_slots, first, steady, n_first, e->tw, e->ybuf, e->Rd, wpos, e->n_streams, tc ? nullptr : nyq_part,
        (uint64_t)e->max_slots * M, (uint64_t)M, 0u, T);
    return;
  }
  k_irfft<M><<<dim3(ceil_div(e->n_streams, FPB), T), dim3(FftCfg<M>::NT, FPB), smem, st>>>(
      ypart, e->max_slots, first, steady, n_first, e->tw, e->ybuf, e->Rd, wpos, e->n_streams,
      tc ? nullptr : nyq_part, (uint64_t)e->max_slots * M, (uint64_t)M, 0u);
}

int launch_irfft(bbx_engine* e, uint32_t T, uint32_t n_first, bool tc, cudaStream_t st) {
  switch (e->B) {
    case 64: launch_irfft_t<64>(e, T, n_first, tc, st); break;
    case 128: launch_irfft_t<128>(e, T, n_first, tc, st); break;
    case 256: launch_irfft_t<256>(e, T, n_first, tc, st); break;
    case 512: launch_irfft_t<512>(e, T, n_first, tc, st); break;
    case 1024: launch_irfft_t<1024>(e, T, n_first, tc, st); break;
    case 2048: launch_irfft_t<2048>(e, T, n_first, tc, st); break;
    case 4096: launch_irfft_t<4096>(e, T, n_first, tc, st); break;
    default: set_error("unsupported block size %u", e->B); return BBX_ERR_UNSUPPORTED;
  }
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

template <int THREADS>
void launch_mac_t(bbx_engine* e, const MacPlan& pl, uint32_t t0, uint32_t nt, cudaStream_t st) {
  const uint32_t halfB = e->B / 2;
  dim3 grid(pl.n_ctas, halfB / THREADS, nt);
  float4* yp = (float4*)e->ypart + (uint64_t)t0 * e->max_slots * halfB;
  float* nq = e->nyq_part + (uint64_t)t0 * e->max_slots;
  const MacSeg* segs = pl.segs();
  const uint32_t* cta = pl.cta_seg_begin();
  const float4* fdl = (const float4*)e->fdl;
  const float keep = e->mac_l2_keep;
  // FDL rows are read once per step only when every input feeds one path; shared inputs (ROUTED fan-out, MIMO)
  // re-read them within the step and must keep normal L2 priority
  const int px = (e->mode == BBX_MODE_PER_CHANNEL) ? 1 : 0;
#define BBX_MAC_LAUNCH(U, OCC)                                                                                            \
  do {                                                                                                                    \
    if (keep > 0.f)                                                                                                       \
      k_fdl_mac<U, THREADS, OCC, true><<<grid, THREADS, 0, st>>>(segs, cta, fdl, yp, nq, halfB, e->R, e->head, t0, e->max_slots, keep, px); \
    else                                                                                                                  \
      k_fdl_mac<U, THREADS, OCC, false><<<grid, THREADS, 0, st>>>(segs, cta, fdl, yp, nq, halfB, e->R, e->head, t0, e->max_slots, keep, px); \
  } while (0)
  // resident CTAs per SM <-> loads in flight per thread: fewer, fatter CTAs unroll deeper
  switch (pl.occ) {
    case 1: BBX_MAC_LAUNCH(16, 1); break;
    case 2: BBX_MAC_LAUNCH(8, 2); break;
    case 3: BBX_MAC_LAUNCH(6, 3); break;
    default: BBX_MAC_LAUNCH(4, 4); break;
  }
#undef BBX_MAC_LAUNCH
}

template <int TT, int THREADS>
void launch_mac_tb_t(bbx_engine* e, const MacPlan& pl, uint32_t t0, uint32_t nt, cudaStream_t st) {
  const uint32_t ncol = e->B / THREADS, ntile = ceil_div(nt, TT);
  dim3 grid(ncol * ntile, pl.n_ctas);
  float2* yp = e->ypart + (uint64_t)t0 * e->max_slots * e->B;
  k_fdl_mac_tb<TT, THREADS, 8><<<grid, THREADS, 0, st>>>(pl.segs(), pl.cta_seg_begin(), e->fdl, yp, e->B, e->R, e->head, t0, nt,
                                                        ncol, e->max_slots);

}

template <int TT>
void launch_mac_tb(bbx_engine* e, const MacPlan& pl, uint32_t t0, uint32_t nt, cudaStream_t st) {
  if (e->B >= 256) launch_mac_tb_t<TT, 256>(e, pl, t0, nt, st);
  else if (e->B == 128) launch_mac_tb_t<TT, 128>(e, pl, t0, nt, st);
  else launch_mac_tb_t<TT, 64>(e, pl, t0, nt, st);
}

// the time-batched kernel pays a window fill of TT-1 rows per term: only worth it for long filters and enough block-steps
bool mac_uses_time_batching(const bbx_engine* e, const MacPlan& pl, uint32_t nt) {
  const uint32_t tb = e->mac_time_tile == 116 ? 16 : e->mac_time_tile;  // 0: streaming only
  return tb && nt >= tb / 2 && pl.n_terms && pl.total_rows / pl.n_terms >= 2 * tb;
}

int launch_mac(bbx_engine* e, const MacPlan& pl, uint32_t t0, uint32_t nt) {
  if (pl.n_ctas == 0 || nt == 0) return BBX_OK;
  cudaStream_t st = e->stream;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e->profile_mac) {
    if (e->mac_events_used + 2 > e->mac_events.size()) {
      size_t old = e->mac_events.size();
      e->mac_events.resize(old + 64);
      for (size_t i = old; i < e->mac_events.size(); i++) BBX_CUDA_TRY(cudaEventCreate(&e->mac_events[i]));
    }
    ev0 = e->mac_events[e->mac_events_used++];
    ev1 = e->mac_events[e->mac_events_used++];
  }
  const uint32_t halfB = e->B / 2;
  const uint32_t tb = e->mac_time_tile;  // 0: streaming only
  const bool use_tb = mac_uses_time_batching(e, pl, nt);
  // k_fdl_mac_tbs covers calls of more than 16 blocks (two or four time tiles per CTA); shorter batched calls and the
  // A/B settings run round 1's per-thread-copy kernel
  const bool use_tbs = use_tb && tb == 16 && nt > 16;
  if (use_tbs) {
    MacTbsArgs a;
    a.segs = pl.segs();
    a.cta_seg_begin = pl.cta_seg_begin();
    a.n_plan_ctas = pl.n_ctas;
    a.fdl = e->fdl;
    a.ypart = e->ypart + (uint64_t)t0 * e->max_slots * e->B;
    a.nyq_part = e->nyq_part + (uint64_t)t0 * e->max_slots;
    a.B = e->B;
    a.R = e->R;
    a.head = e->head;
    a.t0 = t0;
    a.nt = nt;
    a.slot_stride = e->max_slots;
    a.status = e->mac_status;
    // the Nyquist sums of column 0 run next to the MAC on the side stream.  The MAC is enqueued FIRST, on the stream of
    // higher priority, so that its persistent CTAs (two per SM, all of the shared memory) usually take their places before
    // the 148 small CTAs of the side kernel arrive.  Enqueued the other way round, the MAC launch took 0.202 instead of
    // 0.190 ms in four runs of six (same box, same binary): the difference is the side kernel's duration.  An SM that a side
    // CTA reaches first keeps that CTA's shared-memory carve-out until it is empty; the side kernels therefore ask for the
    // MAC's carve-out (launch_nyq_mac2), so that MAC CTAs can join them instead of waiting.
    BBX_CUDA_TRY(cudaEventRecord(e->ev_fork, st));
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_aux, e->ev_fork, 0));
    if (ev0) BBX_CUDA_TRY(cudaEventRecord(ev0, st));
    BBX_CUDA_TRY(launch_mac_tbs(a, st, &e->last_mac_kernel));
    if (ev1) BBX_CUDA_TRY(cudaEventRecord(ev1, st));
    BBX_CUDA_TRY(launch_nyq_mac2(a, e->s_aux));
    BBX_CUDA_TRY(cudaEventRecord(e->ev_join, e->s_aux));
    BBX_CUDA_TRY(cudaStreamWaitEvent(st, e->ev_join, 0));
    e->launches += 2;
  } else {
    if (use_tb) {
      BBX_CUDA_TRY(cudaEventRecord(e->ev_fork, st));
      BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_aux, e->ev_fork, 0));
    }
    if (ev0) BBX_CUDA_TRY(cudaEventRecord(ev0, st));
    e->last_mac_kernel = use_tb ? (tb == 32 ? "k_fdl_mac_tb<32,256,8>" : "k_fdl_mac_tb<16,256,8>") : "k_fdl_mac";
    if (use_tb) {
      if (tb == 32) launch_mac_tb<32>(e, pl, t0, nt, st);
      else launch_mac_tb<16>(e, pl, t0, nt, st);
    } else if (halfB >= 256) launch_mac_t<256>(e, pl, t0, nt, st);
    else if (halfB == 128) launch_mac_t<128>(e, pl, t0, nt, st);
    else if (halfB == 64) launch_mac_t<64>(e, pl, t0, nt, st);
    else launch_mac_t<32>(e, pl, t0, nt, st);
    BBX_CUDA_TRY(cudaGetLastError());
    if (ev1) BBX_CUDA_TRY(cudaEventRecord(ev1, st));
    e->launches++;
    if (use_tb) {
      // Nyquist sums of column 0 (the streaming kernel accumulates them inline): a few hundred latency-bound warps on the
      // side stream, enqueued after the MAC so that they run underneath it without taking its CTAs' places (see above)
      {
        static uint32_t carve_set = 0;  // per device; same reason as in launch_nyq_mac2 (mac_tbs.cu)
        if (!(carve_set & (1u << (e->device & 31)))) {
          cudaFuncSetAttribute(k_nyq_mac, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          carve_set |= 1u << (e->device & 31);
        }
      }
      k_nyq_mac<<<dim3(pl.n_ctas, ceil_div(nt, 32)), 32, 0, e->s_aux>>>(pl.segs(), pl.cta_seg_begin(), e->fdl,
                                                                       e->nyq_part + (uint64_t)t0 * e->max_slots, e->B, e->R,
                                                                       e->head, t0, nt, e->max_slots);
      BBX_CUDA_TRY(cudaGetLastError());
      BBX_CUDA_TRY(cudaEventRecord(e->ev_join, e->s_aux));
      e->launches++;
      BBX_CUDA_TRY(cudaStreamWaitEvent(st, e->ev_join, 0));
    }
  }
  e->mac_launches++;
  e->mac_units += (uint64_t)e->n_streams * nt;
  // 16 P K + 16 K + (bytes_in + bytes_out) B per channel-block, K = B + 1
  const uint64_t K = e->B + 1;
  e->mac_bytes += (uint64_t)nt * (16ull * pl.total_rows * K + 16ull * K * e->n_streams +
                                  (uint64_t)e->B * (fmt_bytes(e->last_infmt) * e->n_in + fmt_bytes(e->last_outfmt) * e->n_out));
  return BBX_OK;
}

// The pinned plan/route staging is rewritten by the host; wait until the previous upload has read it.
int wait_uploads(bbx_engine* e) {
  if (e->upload_pending) {
    BBX_CUDA_TRY(cudaEventSynchronize(e->ev_upload));
    e->upload_pending = false;
  }
  return BBX_OK;
}
int mark_upload(bbx_engine* e) {
  BBX_CUDA_TRY(cudaEventRecord(e->ev_upload, e->stream));
  e->upload_pending = true;
  return BBX_OK;
}

int staging_alloc(Staging& s, size_t bytes) {
  s.bytes = bytes;
  for (int i = 0; i < Staging::kDepth; i++) {
    BBX_CUDA_TRY(cudaHostAlloc((void**)&s.h[i], bytes, cudaHostAllocDefault));
    memset(s.h[i], 0, bytes);
    BBX_CUDA_TRY(cudaEventCreateWithFlags(&s.ev[i], cudaEventDisableTiming));
  }
  return BBX_OK;
}
void staging_free(Staging& s) {
  for (int i = 0; i < Staging::kDepth; i++) {
    if (s.h[i]) cudaFreeHost(s.h[i]);
    if (s.ev[i]) cudaEventDestroy(s.ev[i]);
    s.h[i] = nullptr;
    s.ev[i] = nullptr;
  }
}
// next slot to fill; blocks only while the copy that last read this slot is still queued
int staging_acquire(Staging& s, uint8_t** out) {
  if (s.pending[s.cur]) {
    BBX_CUDA_TRY(cudaEventSynchronize(s.ev[s.cur]));
    s.pending[s.cur] = false;
  }
  *out = s.h[s.cur];
  return BBX_OK;
}
int staging_commit(Staging& s, void* dst, cudaStream_t st) {
  BBX_CUDA_TRY(cudaMemcpyAsync(dst, s.h[s.cur], s.bytes, cudaMemcpyHostToDevice, st));
  BBX_CUDA_TRY(cudaEventRecord(s.ev[s.cur], st));
  s.pending[s.cur] = true;
  s.cur = (s.cur + 1) % Staging::kDepth;
  return BBX_OK;
}

// Build a MAC plan from a job list into a pinned staging slot and enqueue its upload.
int build_plan(bbx_engine* e, MacPlan& pl, const std::vector<std::vector<JobTerm>>& jobs, const std::vector<uint32_t>& xjob) {
  uint32_t total = 0, nterms = 0;
  for (auto& j : jobs)
    for (auto& tm : j)
      if (tm.f) {
        total += tm.f->P;
        nterms++;
      }
  pl.n_terms = nterms;
  {
    int wrc = staging_acquire(pl.stg, &pl.h_blob);
    if (wrc) return wrc;
  }
  MacSeg* segs = (MacSeg*)(pl.h_blob + pl.off_segs);
  uint32_t* cta = (uint32_t*)(pl.h_blob + pl.off_cta);
  uint32_t* jfirst = (uint32_t*)(pl.h_blob + pl.off_first);
  uint32_t* jcount = (uint32_t*)(pl.h_blob + pl.off_count);
  uint32_t* xj = (uint32_t*)(pl.h_blob + pl.off_xjob);
  uint32_t* jseg = (uint32_t*)(pl.h_blob + pl.off_jseg);
  BBX_REQUIRE(jobs.size() <= e->max_jobs, "internal: too many jobs");
  pl.max_job_rows = 0;
  for (auto& j : jobs) {
    uint32_t rows = 0;
    for (auto& tm : j)
      if (tm.f) rows += tm.f->P;
    pl.max_job_rows = std::max(pl.max_job_rows, rows);
  }
  pl.n_jobs = (uint32_t)jobs.size();
  pl.total_rows = total;
  for (uint32_t s = 0; s < e->n_streams; s++) xj[s] = s < xjob.size() ? xjob[s] : kNoJob;
  if (total == 0) {
    for (uint32_t j = 0; j < jobs.size(); j++) jfirst[j] = jcount[j] = 0;
    for (uint32_t j = 0; j <= jobs.size(); j++) jseg[j] = 0;
    pl.n_ctas = 0;
    pl.n_slots = 0;
  } else {
    // even split of the flattened row space; small problems get fewer, fatter CTAs
    const uint32_t min_rows = 4;
    // short filters (<= 8 partitions per term, e.g. the 64x64 MIMO matrix) cannot fill a 16-deep unroll: cut the
    // plan for two resident CTAs per SM with 8 rows in flight each instead
    pl.occ = (nterms && total / nterms <= 8 && e->mac_occ < 2) ? 2u : e->mac_occ;
    uint32_t G = std::min(kNumSMs * pl.occ, std::max(1u, total / min_rows));
    uint32_t rpc = ceil_div(total, G);
    G = ceil_div(total, rpc);
    uint32_t nseg = 0, nslot = 0, row = 0, cur_cta = 0;
    cta[0] = 0;
    int run_job = -1;  // job of the open run inside the current CTA
    for (uint32_t j = 0; j < jobs.size(); j++) {
      jfirst[j] = nslot;
      jseg[j] = nseg;  // a job's segments are consecutive: the fused single-launch path walks them in this order
      uint32_t before = nslot;
      for (auto& tm : jobs[j]) {
        if (!tm.f) continue;
        uint32_t p = 0;
        while (p < tm.f->P) {
          uint32_t cta_end = (cur_cta + 1) * rpc;
          if (row == cta_end) {  // move to the next CTA
            cur_cta++;
            cta[cur_cta] = nseg;
            run_job = -1;
            continue;
          }
          uint32_t np = std::min(tm.f->P - p, cta_end - row);
          BBX_REQUIRE(nseg < e->max_segs, "internal: MAC plan overflow (segments)");
          MacSeg& sg = segs[nseg];
          sg.H = (const float4*)tm.f->H;
          sg.fdl_ch = tm.input;
          sg.p0 = p;
          sg.np = np;
          sg.pad = 0;
          if (run_job == (int)j) {
            // continue the open run: the previous segment no longer writes
            segs[nseg - 1].flags &= ~2u;
            sg.flags = 2u;
            sg.slot = segs[nseg - 1].slot;
          } else {
            BBX_REQUIRE(nslot < e->max_slots, "internal: MAC plan overflow (slots)");
            sg.flags = 1u | 2u;
            sg.slot = nslot++;
            run_job = (int)j;
          }
          nseg++;
          p += np;
          row += np;
        }
      }
      jcount[j] = nslot - before;
    }
    jseg[jobs.size()] = nseg;
    for (uint32_t c = cur_cta + 1; c <= G; c++) cta[c] = nseg;
    pl.n_ctas = G;
    pl.n_slots = nslot;
  }
  pl.valid = true;
  return staging_commit(pl.stg, pl.d_blob, e->stream);
}

int alloc_plan(bbx_engine* e, MacPlan& pl) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  pl.off_segs = take(sizeof(MacSeg) * e->max_segs);
  pl.off_cta = take(sizeof(uint32_t) * (e->max_ctas + 2));
  pl.off_first = take(sizeof(uint32_t) * e->max_jobs);
  pl.off_count = take(sizeof(uint32_t) * e->max_jobs);
  pl.off_xjob = take(sizeof(uint32_t) * std::max(1u, e->n_streams));
  pl.off_jseg = take(sizeof(uint32_t) * (e->max_jobs + 1));
  pl.blob_bytes = off;
  {
    int src = staging_alloc(pl.stg, off);
    if (src) return src;
  }
  BBX_CUDA_TRY(cudaMalloc((void**)&pl.d_blob, off));
  BBX_CUDA_TRY(cudaMemset(pl.d_blob, 0, off));
  return BBX_OK;
}

// jobs for the given choice of filter per path; MIMO groups the paths of one output into one job
void make_jobs(const bbx_engine* e, bool use_pending, std::vector<std::vector<JobTerm>>& jobs) {
  jobs.clear();
  auto pick = [&](const PathState& p) { return (use_pending && p.has_pending) ? p.pend : p.cur; };
  if (e->mode == BBX_MODE_MIMO) {
    jobs.resize(e->n_out);
    for (uint32_t o = 0; o < e->n_out; o++)
      for (uint32_t i = 0; i < e->n_in; i++) {
        const PathState& p = e->paths[(size_t)o * e->n_in + i];
        const bbx_filter* f = pick(p);
        if (f) jobs[o].push_back({f, i});
      }
  } else {
    jobs.resize(e->n_paths);
    for (uint32_t k = 0; k < e->n_paths; k++) {
      const PathState& p = e->paths[k];
      const bbx_filter* f = pick(p);
      if (f) jobs[k].push_back({f, p.input});
    }
  }
}

int upload_routes(bbx_engine* e, bool first_block_transition) {
  {
    int wrc = staging_acquire(e->route_stg, &e->h_route);
    if (wrc) return wrc;
  }
  uint32_t* ofirst = (uint32_t*)(e->h_route + e->roff_first);
  uint32_t* rstream = (uint32_t*)(e->h_route + e->roff_stream);
  float* gain = (float*)(e->h_route + e->roff_gain);
  double* dcur = (double*)(e->h_route + e->roff_dcur);
  double* dold = (double*)(e->h_route + e->roff_dold);
  uint32_t* flags = (uint32_t*)(e->h_route + e->roff_flags);
  uint32_t* icur = (uint32_t*)(e->h_route + e->roff_icur);
  uint32_t* iold = (uint32_t*)(e->h_route + e->roff_iold);
  if (e->mode == BBX_MODE_MIMO) {
    for (uint32_t o = 0; o < e->n_out; o++) {
      ofirst[o] = o;
      rstream[o] = o;
      gain[o] = 1.0f;
      dcur[o] = dold[o] = 0.0;
      icur[o] = iold[o] = 0;
      flags[o] = 0;
    }
    ofirst[e->n_out] = e->n_out;
    // sharded: PCM channel c of this rank is output sh_o0 + c
    for (uint32_t c = 0; c < e->n_out_pcm && e->n_out_pcm < e->n_out; c++) rstream[c] = e->sh_o0 + c;
  } else {
    uint32_t n = 0;
    for (uint32_t o = 0; o < e->n_out; o++) {
      ofirst[o] = n;
      for (uint32_t k = 0; k < e->n_paths; k++)
        if (e->paths[k].output == o) rstream[n++] = k;
    }
    ofirst[e->n_out] = n;
    for (uint32_t k = 0; k < e->n_paths; k++) {
      const PathState& p = e->paths[k];
      gain[k] = p.gain;
      bool sw = first_block_transition && p.has_pending;
      dcur[k] = sw ? p.pend_delay : p.delay;
      dold[k] = p.delay;
      icur[k] = (uint32_t)dcur[k] % e->Rd;
      iold[k] = (uint32_t)dold[k] % e->Rd;
      flags[k] = (sw && p.xfade && p.pend_delay != p.delay) ? 1u : 0u;
    }
  }
  // stream -> input (what the fused single-launch kernel reads; MIMO streams have many inputs and never take that path)
  {
    uint32_t* sin = (uint32_t*)(e->h_route + e->roff_input);
    for (uint32_t k = 0; k < e->n_streams; k++) sin[k] = (e->mode == BBX_MODE_MIMO) ? 0u : e->paths[k].input;
  }
  // per-route entries in mixdown order (what k_pcm_out reads)
  {
    RouteEntry* en = (RouteEntry*)(e->h_route + e->roff_entry);
    const uint32_t nroutes = ofirst[e->n_out_pcm < e->n_out ? e->n_out_pcm : e->n_out];
    e->n_routes_pcm = nroutes;
    for (uint32_t r = 0; r < nroutes; r++) {
      const uint32_t st = rstream[r];
      en[r].stream = st;
      en[r].gain = gain[st];
      en[r].icur = icur[st];
      en[r].iold = iold[st];
      en[r].flags = flags[st];
      en[r].pad = 0;
      en[r].dcur = dcur[st];
      en[r].dold = dold[st];
    }
  }
  return staging_commit(e->route_stg, e->d_route, e->stream);
}

RouteView route_view(const bbx_engine* e) {
  RouteView v;
  v.out_first = (const uint32_t*)(e->d_route + e->roff_first);
  v.entry = (const RouteEntry*)(e->d_route + e->roff_entry);
  return v;
}

bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

// ---- MIMO on the tensor cores ----
int tc_alloc(bbx_engine* e) {
  uint32_t P2 = 1, lg = 0;
  while (P2 < e->Pmax) {
    P2 <<= 1;
    lg++;
  }
  e->tc_P2 = P2;
  e->tc_P2log = lg;
  const uint32_t Kc = ceil_div(e->n_in * P2, (uint32_t)kTcChunk) * kTcChunk;  // complex K, padded to whole chunks
  e->tc_G = Kc / 2;
  e->tc_nog = ceil_div(e->n_out, 64u);
  // pre-staged FDL runs: per bin [column tile][chunk][hi | lo][rows of the chunk][seg], sized for N = 64
  {
    const uint32_t ninp = lg < 4 ? (16u >> lg) : 1u, seg = lg < 4 ? ((kTcNmax + P2) & ~1u) : kTcNmax + 16;
    e->tc_xbin_max = (uint64_t)ceil_div(e->Tmax, (uint32_t)kTcNmax) * (Kc / kTcChunk) * 2 * ninp * seg;
  }
  const size_t hbytes = sizeof(float4) * (size_t)e->tc_nog * e->B * e->tc_G * 64;
  const size_t xbytes = sizeof(float2) * (size_t)e->B * e->tc_xbin_max;
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_hpack, hbytes));
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_xb, xbytes));
  const size_t npaths = (size_t)e->n_out * e->n_in;
  BBX_CUDA_TRY(cudaHostAlloc((void**)&e->tc_ftab_h, sizeof(float2*) * npaths, cudaHostAllocDefault));
  BBX_CUDA_TRY(cudaHostAlloc((void**)&e->tc_fparts_h, sizeof(uint32_t) * npaths, cudaHostAllocDefault));
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_ftab_d, sizeof(float2*) * npaths));
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_fparts_d, sizeof(uint32_t) * npaths));
  std::vector<uint32_t> view(3 * (size_t)e->n_out);
  for (uint32_t o = 0; o < e->n_out; o++) {
    view[o] = o;
    view[e->n_out + o] = 1;
    view[2 * (size_t)e->n_out + o] = kNoJob;
  }
  BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_view, sizeof(uint32_t) * view.size()));
  BBX_CUDA_TRY(cudaMemcpy(e->tc_view, view.data(), sizeof(uint32_t) * view.size(), cudaMemcpyHostToDevice));
  BBX_CUDA_TRY(cudaHostAlloc((void**)&e->tc_status_h, sizeof(int), cudaHostAllocMapped));
  *e->tc_status_h = 0;
  BBX_CUDA_TRY(cudaHostGetDevicePointer((void**)&e->tc_status, e->tc_status_h, 0));
  BBX_CUDA_TRY(cudaFuncSetAttribute(k_mimo_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemMax));
  BBX_CUDA_TRY(cudaFuncSetAttribute(k_mimo_tc<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemMax));
  BBX_CUDA_TRY(cudaFuncSetAttribute(k_mimo_tc<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemMax));
  e->tc_on = true;
  e->tc_dirty = true;
  return BBX_OK;
}

// rebuild the bin-major filter operand from the current filter matrix
int tc_pack_filters(bbx_engine* e) {
  int rc = wait_uploads(e);
  if (rc) return rc;
  const size_t npaths = (size_t)e->n_out * e->n_in;
  for (size_t k = 0; k < npaths; k++) {
    const bbx_filter* f = e->paths[k].cur;
    e->tc_ftab_h[k] = f ? f->H : nullptr;
    e->tc_fparts_h[k] = f ? f->P : 0;
  }
  BBX_CUDA_TRY(cudaMemcpyAsync(e->tc_ftab_d, e->tc_ftab_h, sizeof(float2*) * npaths, cudaMemcpyHostToDevice, e->stream));
  BBX_CUDA_TRY(cudaMemcpyAsync(e->tc_fparts_d, e->tc_fparts_h, sizeof(uint32_t) * npaths, cudaMemcpyHostToDevice, e->stream));
  if ((rc = mark_upload(e))) return rc;
  k_mimo_pack_h<<<dim3(e->B / 32, e->tc_G / 2, e->tc_nog), 256, 0, e->stream>>>(e->tc_ftab_d, e->tc_fparts_d, e->tc_hpack, e->B, e->n_in,
                                                                        e->n_out, e->tc_P2log, e->tc_G);
  BBX_CUDA_TRY(cudaGetLastError());
  e->launches++;
  e->tc_dirty = false;
  return BBX_OK;
}

// the T block-steps of one call: FDL rows -> bin-major X, then one GEMM per bin
int launch_mimo_tc(bbx_engine* e, uint32_t T) {
  cudaStream_t st = e->stream;
  uint32_t N = 16, Nlog = 4;
  while (N < T && N < (uint32_t)kTcNmax) {
    N <<= 1;
    Nlog++;
  }
  const uint32_t ntiles = ceil_div(T, N);
  const uint32_t nchunk = e->tc_G / (kTcChunk / 2);
  const uint32_t ninp = e->tc_P2log < 4 ? (16u >> e->tc_P2log) : 1u;
  const uint32_t seg = e->tc_P2log < 4 ? ((N + e->tc_P2) & ~1u) : N + 16;
  const uint64_t xbin = (uint64_t)ntiles * nchunk * 2 * ninp * seg;
  k_mimo_pack_x<<<dim3(e->B / 32, ntiles * nchunk * ninp, ceil_div(seg, 32u)), 256, 0, st>>>(e->fdl, e->tc_xb, e->B, e->R, e->head,
                                                                                         e->n_in, e->tc_P2log, T, N, nchunk, xbin);
  BBX_CUDA_TRY(cudaGetLastError());
  e->launches++;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e->profile_mac) {
    if (e->mac_events_used + 2 > e->mac_events.size()) {
      size_t old = e->mac_events.size();
      e->mac_events.resize(old + 64);
      for (size_t i = old; i < e->mac_events.size(); i++) BBX_CUDA_TRY(cudaEventCreate(&e->mac_events[i]));
    }
    ev0 = e->mac_events[e->mac_events_used++];
    ev1 = e->mac_events[e->mac_events_used++];
  }
  MimoTcArgs a;
  a.hpack = e->tc_hpack;
  a.xb = e->tc_xb;
  a.ypart = e->ypart;
  a.status = e->tc_status;
  a.B = e->B;
  a.n_in = e->n_in;
  a.n_out = e->n_out;
  a.P2log = e->tc_P2log;
  a.G = e->tc_G;
  a.T = T;
  a.slot_stride = e->max_slots;
  a.xbin = xbin;
  a.trace = nullptr;
  if (ev0) BBX_CUDA_TRY(cudaEventRecord(ev0, st));
  // raw ring: the hi and lo FDL runs of a chunk (see mimo_tc.cuh); as many stages as fit, an even number
  {
    a.raw_stage_bytes = (2 * ninp * seg * 8 + 127) & ~127u;
    uint32_t nst = (kTcSmemMax - kTcOffRaw) / a.raw_stage_bytes;
    nst = std::min(nst, (uint32_t)kTcRawStagesMax) & ~1u;
    a.raw_stages = nst;
  }
  const uint32_t smem = kTcOffRaw + a.raw_stages * a.raw_stage_bytes;
  const dim3 grid(e->B / kTcBins, e->tc_nog, ntiles);
  if (e->tc_trace && grid.x * grid.y * grid.z <= e->tc_trace_ctas) a.trace = e->tc_trace;
  if (Nlog == 4) k_mimo_tc<4><<<grid, kTcThreads, smem, st>>>(a);
  else if (Nlog == 5) k_mimo_tc<5><<<grid, kTcThreads, smem, st>>>(a);
  else k_mimo_tc<6><<<grid, kTcThreads, smem, st>>>(a);
  BBX_CUDA_TRY(cudaGetLastError());
  if (ev1) BBX_CUDA_TRY(cudaEventRecord(ev1, st));
  e->launches++;
  e->last_mac_kernel = "k_mimo_tc";
  e->tc_launches++;
  e->mac_launches++;
  e->mac_units += (uint64_t)e->n_streams * T;
  const uint64_t K = e->B + 1;
  e->mac_bytes += (uint64_t)T * (16ull * e->plan_steady.total_rows * K + 16ull * K * e->n_streams +
                                 (uint64_t)e->B * (fmt_bytes(e->last_infmt) * e->n_in + fmt_bytes(e->last_outfmt) * e->n_out));
  return BBX_OK;
}

template <int M, bool FUSE_OUT>
void launch_fused_t(const FusedArgs& a, cudaStream_t st) {
  constexpr int FPB = FftCfg<M>::FPB;
  constexpr size_t smem = sizeof(float2) * (size_t)FPB * (M + FftCfg<M>::MP);
  if (smem > 48 * 1024) {
    static uint32_t attr_set = 0;  // per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_set & (1u << (dev & 31)))) {
      cudaFuncSetAttribute(k_block_fused<M, FUSE_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr_set |= 1u << (dev & 31);
    }
  }
  k_block_fused<M, FUSE_OUT><<<ceil_div(a.n_streams, (uint32_t)FPB), dim3(FftCfg<M>::NT, FPB), smem, st>>>(a);
}
template <bool FUSE_OUT>
int launch_fused(uint32_t B, const FusedArgs& a, cudaStream_t st) {
  switch (B) {
    case 64: launch_fused_t<64, FUSE_OUT>(a, st); break;
    case 128: launch_fused_t<128, FUSE_OUT>(a, st); break;
    case 256: launch_fused_t<256, FUSE_OUT>(a, st); break;
    case 512: launch_fused_t<512, FUSE_OUT>(a, st); break;
    case 1024: launch_fused_t<1024, FUSE_OUT>(a, st); break;
    case 2048: launch_fused_t<2048, FUSE_OUT>(a, st); break;
    case 4096: launch_fused_t<4096, FUSE_OUT>(a, st); break;
    default: set_error("unsupported block size %u", B); return BBX_ERR_UNSUPPORTED;
  }
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

// 0 always means "leave as is" (library default at creation)
void apply_tuning(bbx_engine* e, uint32_t ctas_per_sm, uint32_t l2_keep_16ths, uint32_t time_tile) {
  if (ctas_per_sm) e->mac_occ = std::min(ctas_per_sm, 4u);
  if (l2_keep_16ths) e->mac_l2_keep = (l2_keep_16ths > 16u) ? 0.f : l2_keep_16ths / 16.0f;  // > 16: hints off
  if (time_tile) e->mac_time_tile = (time_tile == 16 || time_tile == 32 || time_tile == 116) ? time_tile : 0;  // 1: streaming only
  e->steady_dirty = true;
}

}  // namespace

extern "C" {

int bbx_engine_create(const bbx_config* cfg, bbx_engine** out) {
  BBX_REQUIRE(cfg && out, "bbx_engine_create: null argument");
  int rc = require_device();
  if (rc) return rc;
  BBX_REQUIRE(is_pow2(cfg->block_size) && cfg->block_size >= 64 && cfg->block_size <= 4096,
              "block_size %u must be a power of two in [64, 4096]", cfg->block_size);
  BBX_REQUIRE(cfg->n_inputs > 0, "n_inputs must be > 0");
  BBX_REQUIRE(cfg->mode >= BBX_MODE_PER_CHANNEL && cfg->mode <= BBX_MODE_MIMO, "bad mode %d", cfg->mode);
  bbx_engine* e = new bbx_engine();
  e->cfg = *cfg;
  e->device = cfg->device;
  // everything below returns through BBX_REQUIRE / BBX_CUDA_TRY: run it as one unit so that a failure half way releases
  // what was allocated so far (bbx_engine_destroy accepts a partially built engine)
  DeviceGuard dg(e->device);  // the caller's current device comes back on return
  rc = [&]() -> int {
  {
    int cur = -1;
    BBX_CUDA_TRY(cudaGetDevice(&cur));
    BBX_REQUIRE(cur == e->device, "bbx_engine_create: device %d cannot be selected", e->device);
  }
  e->B = cfg->block_size;
  e->Pmax = std::max(1u, cfg->max_partitions);
  e->n_in = cfg->n_inputs;
  e->mode = cfg->mode;
  if (e->mode == BBX_MODE_PER_CHANNEL) {
    e->n_out = e->n_in;
    e->n_paths = e->n_in;
    e->n_streams = e->n_in;
  } else if (e->mode == BBX_MODE_ROUTED) {
    BBX_REQUIRE(cfg->n_outputs > 0 && cfg->n_paths > 0, "ROUTED mode needs n_outputs and n_paths");
    e->n_out = cfg->n_outputs;
    e->n_paths = cfg->n_paths;
    e->n_streams = cfg->n_paths;
  } else {
    BBX_REQUIRE(cfg->n_outputs > 0, "MIMO mode needs n_outputs");
    BBX_REQUIRE(cfg->max_delay == 0, "MIMO mode has no per-path delay (frequency-domain mixdown)");
    e->n_out = cfg->n_outputs;
    e->n_paths = e->n_in * e->n_out;
    e->n_streams = e->n_out;
    if (cfg->mimo_shard_world > 1) {
      BBX_REQUIRE(cfg->mimo_shard_rank < cfg->mimo_shard_world, "mimo_shard_rank %u out of range", cfg->mimo_shard_rank);
      BBX_REQUIRE(e->n_out % cfg->mimo_shard_world == 0, "n_outputs %u is not a multiple of mimo_shard_world %u", e->n_out,
                  cfg->mimo_shard_world);
      e->sh_world = cfg->mimo_shard_world;
      e->sh_rank = cfg->mimo_shard_rank;
    }
  }
  BBX_REQUIRE(e->mode == BBX_MODE_MIMO || cfg->mimo_shard_world <= 1, "mimo_shard_world needs MIMO mode");
  e->sh_nloc = e->n_out / e->sh_world;
  e->sh_o0 = e->sh_rank * e->sh_nloc;
  e->n_out_pcm = (e->sh_world > 1) ? e->sh_nloc : e->n_out;
  e->Tmax = std::max(1u, cfg->max_blocks);
  e->R = e->Pmax + e->Tmax - 1;
  uint32_t need = cfg->max_delay + 14 + (e->Tmax + 1) * e->B;
  e->Rd = cfg->ring_length ? cfg->ring_length : ceil_div(need, e->B) * e->B;
  BBX_REQUIRE(e->Rd >= cfg->max_delay + 14 + e->Tmax * e->B, "ring_length %u too short (need >= %u)", e->Rd,
              cfg->max_delay + 14 + e->Tmax * e->B);
  e->xstride = (e->Tmax + 1) * e->B;
  e->paths.resize(e->n_paths);
  for (uint32_t k = 0; k < e->n_paths; k++) {
    PathState& p = e->paths[k];
    if (e->mode == BBX_MODE_MIMO) {
      p.output = k / e->n_in;
      p.input = k % e->n_in;
    } else if (e->mode == BBX_MODE_PER_CHANNEL) {
      p.input = p.output = k;
    }
  }
  {
    // the engine stream outranks the side stream: when the MAC and its side kernel become ready together, the MAC's persistent
    // CTAs are placed first and the side kernel fills what is left (launch_mac: the enqueue order alone did not hold in every
    // process -- a 2-rank run still showed the MAC 12 us longer, the side kernel's duration)
    int prio_low = 0, prio_high = 0;
    BBX_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
    BBX_CUDA_TRY(cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_high));
    BBX_CUDA_TRY(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    BBX_CUDA_TRY(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
    BBX_CUDA_TRY(cudaStreamCreateWithPriority(&e->s_aux, cudaStreamNonBlocking, prio_low));
  }
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  for (int i = 0; i < bbx_engine::kIoSlots; i++) {
    BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming));
    BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_comp[i], cudaEventDisableTiming));
    BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_d2h[i], cudaEventDisableTiming));
  }
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join_in, cudaEventDisableTiming));
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join_out, cudaEventDisableTiming));
  BBX_CUDA_TRY(cudaEventCreate(&e->ev_start));
  BBX_CUDA_TRY(cudaEventCreate(&e->ev_stop));
  BBX_CUDA_TRY(cudaEventCreateWithFlags(&e->ev_upload, cudaEventDisableTiming));
  BBX_CUDA_TRY(cudaHostAlloc((void**)&e->mac_status_h, sizeof(int), cudaHostAllocMapped));
  *e->mac_status_h = 0;
  BBX_CUDA_TRY(cudaHostGetDevicePointer((void**)&e->mac_status, e->mac_status_h, 0));

  const uint32_t B = e->B, N = 2 * B;
  // twiddles exp(-2 pi i j / N), computed in double
  {
    std::vector<float2> tw(N);
    const double PI = 3.14159265358979323846264338327950288;
    for (uint32_t j = 0; j < N; j++) {
      double a = -2.0 * PI * (double)j / (double)N;
      tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->tw, sizeof(float2) * N));
    BBX_CUDA_TRY(cudaMemcpy(e->tw, tw.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
  }
  size_t xin_bytes = sizeof(float) * (size_t)e->n_in * e->xstride;
  for (int i = 0; i < 2; i++) {
    BBX_CUDA_TRY(cudaMalloc((void**)&e->xin[i], xin_bytes));
    BBX_CUDA_TRY(cudaMemset(e->xin[i], 0, xin_bytes));
  }
  size_t fdl_bytes = sizeof(float2) * (size_t)e->n_in * e->R * B;
  BBX_CUDA_TRY(cudaMalloc((void**)&e->fdl, fdl_bytes));
  BBX_CUDA_TRY(cudaMemset(e->fdl, 0, fdl_bytes));
  size_t ybuf_bytes = sizeof(float) * (size_t)e->n_streams * e->Rd;
  BBX_CUDA_TRY(cudaMalloc((void**)&e->ybuf, ybuf_bytes));
  BBX_CUDA_TRY(cudaMemset(e->ybuf, 0, ybuf_bytes));

  // plan capacities
  e->max_ctas = kNumSMs * 4;  // capacity for every tuning; the plan uses kNumSMs * mac_occ of them
  apply_tuning(e, cfg->mac_ctas_per_sm, cfg->mac_l2_keep_16ths, cfg->mac_time_tile);
  uint32_t terms = (e->mode == BBX_MODE_MIMO) ? e->n_paths : e->n_paths;
  e->max_jobs = 2 * e->n_streams + 1;
  e->max_segs = e->max_ctas + 2 * terms + 8;
  e->max_slots = e->max_ctas + e->max_jobs + 8;
  if ((rc = alloc_plan(e, e->plan_first))) return rc;
  if ((rc = alloc_plan(e, e->plan_steady))) return rc;
  size_t ypart_bytes = sizeof(float2) * (size_t)e->Tmax * e->max_slots * B;
  BBX_CUDA_TRY(cudaMalloc((void**)&e->ypart, ypart_bytes));
  BBX_CUDA_TRY(cudaMemset(e->ypart, 0, ypart_bytes));
  BBX_CUDA_TRY(cudaMalloc((void**)&e->nyq_part, sizeof(float) * (size_t)e->Tmax * e->max_slots));
  BBX_CUDA_TRY(cudaMemset(e->nyq_part, 0, sizeof(float) * (size_t)e->Tmax * e->max_slots));

  if (e->sh_world > 1) {
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_send, sizeof(float2) * (size_t)e->n_out * e->Tmax * B));
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_recv, sizeof(float2) * (size_t)e->sh_nloc * e->Tmax * B));
    std::vector<uint32_t> view(3 * (size_t)e->n_out);
    for (uint32_t o = 0; o < e->n_out; o++) {
      view[o] = (o >= e->sh_o0 && o < e->sh_o0 + e->sh_nloc) ? o - e->sh_o0 : 0u;
      view[e->n_out + o] = 1;
      view[2 * (size_t)e->n_out + o] = kNoJob;
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_view, sizeof(uint32_t) * view.size()));
    BBX_CUDA_TRY(cudaMemcpy(e->sh_view, view.data(), sizeof(uint32_t) * view.size(), cudaMemcpyHostToDevice));
  }

  // MIMO: tensor-core operands (time-batched calls only)
  if (e->mode == BBX_MODE_MIMO && cfg->mimo_tensor != 1 && e->Tmax >= e->tc_min_blocks && e->B >= 64) {
    uint32_t P2 = 1;
    while (P2 < e->Pmax) P2 <<= 1;
    // longer sums than kTcMaxK complex terms stay on the exact-fp32 SIMT MAC (accumulator truncation, mimo_tc.cuh)
    if ((uint64_t)e->n_in * P2 <= kTcMaxK) {
      if ((rc = tc_alloc(e))) return rc;
    }
  }

  // route tables
  {
    size_t off = 0;
    auto take = [&](size_t bytes) {
      size_t o = off;
      off += (bytes + 255) & ~(size_t)255;
      return o;
    };
    uint32_t ns = e->n_streams;
    e->roff_first = take(sizeof(uint32_t) * (e->n_out + 1));
    e->roff_stream = take(sizeof(uint32_t) * ns);
    e->roff_gain = take(sizeof(float) * ns);
    e->roff_dcur = take(sizeof(double) * ns);
    e->roff_dold = take(sizeof(double) * ns);
    e->roff_flags = take(sizeof(uint32_t) * ns);
    e->roff_icur = take(sizeof(uint32_t) * ns);
    e->roff_iold = take(sizeof(uint32_t) * ns);
    e->roff_entry = take(sizeof(RouteEntry) * ns);
    e->roff_input = take(sizeof(uint32_t) * ns);
    e->route_bytes = off;
    {
      int src = staging_alloc(e->route_stg, off);
      if (src) return src;
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->d_route, off));
  }
  BBX_CUDA_TRY(cudaDeviceSynchronize());
  return BBX_OK;
  }();
  if (rc) {
    const std::string msg = get_error();  // the clean-up below must not replace the message of the failure
    bbx_engine_destroy(e);
    set_error("%s", msg.c_str());
    return rc;
  }
  *out = e;
  return BBX_OK;
}

int bbx_engine_destroy(bbx_engine* e) {
  if (!e) return BBX_OK;
  DeviceGuard dg(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->s_in) cudaStreamSynchronize(e->s_in);
  if (e->s_out) cudaStreamSynchronize(e->s_out);
  if (e->s_aux) cudaStreamSynchronize(e->s_aux);
  cudaFree(e->tw);
  cudaFree(e->xin[0]);
  cudaFree(e->xin[1]);
  cudaFree(e->fdl);
  cudaFree(e->ypart);
  cudaFree(e->nyq_part);
  cudaFree(e->ybuf);
  for (cudaEvent_t ev : e->io_ev) cudaEventDestroy(ev);
  for (int i = 0; i < bbx_engine::kIoSlots; i++) {
    cudaFree(e->d_in[i]);
    cudaFree(e->d_out[i]);
    if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]);
    if (e->ev_comp[i]) cudaEventDestroy(e->ev_comp[i]);
    if (e->ev_d2h[i]) cudaEventDestroy(e->ev_d2h[i]);
  }
  if (e->ev_join_in) cudaEventDestroy(e->ev_join_in);
  if (e->ev_join_out) cudaEventDestroy(e->ev_join_out);
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  if (e->s_aux) cudaStreamDestroy(e->s_aux);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  cudaFree(e->flush_buf);
  cudaFree(e->d_lat_in);
  cudaFree(e->d_lat_out);
  for (void* pp : e->px_peer)
    if (pp) cudaIpcCloseMemHandle(pp);
  cudaFree(e->px_mem);
  cudaFree(e->px_done);
  cudaFree(e->px_view);
  if (e->px_status_h) cudaFreeHost(e->px_status_h);
  cudaFree(e->sh_send);
  cudaFree(e->sh_recv);
  cudaFree(e->sh_view);
  cudaFree(e->tc_hpack);
  cudaFree(e->tc_xb);
  cudaFree(e->tc_ftab_d);
  cudaFree(e->tc_fparts_d);
  cudaFree(e->tc_view);
  if (e->tc_status_h) cudaFreeHost(e->tc_status_h);
  if (e->mac_status_h) cudaFreeHost(e->mac_status_h);
  cudaFree(e->tc_trace);
  if (e->tc_ftab_h) cudaFreeHost(e->tc_ftab_h);
  if (e->tc_fparts_h) cudaFreeHost(e->tc_fparts_h);
  cudaFree(e->d_route);
  staging_free(e->route_stg);
  for (MacPlan* pl : {&e->plan_first, &e->plan_steady}) {
    cudaFree(pl->d_blob);
    staging_free(pl->stg);
  }
  for (cudaEvent_t ev : e->mac_events) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->xchg_events) cudaEventDestroy(ev);
  if (e->ev_start) cudaEventDestroy(e->ev_start);
  if (e->ev_stop) cudaEventDestroy(e->ev_stop);
  if (e->ev_upload) cudaEventDestroy(e->ev_upload);
  if (e->stream) cudaStreamDestroy(e->stream);
  // filters that were never destroyed die with their engine (their handles are invalid from here on)
  {
    std::lock_guard<std::mutex> lk(g_filter_mu);
    for (bbx_filter* f : e->filters) {
      g_live_filters.erase(f);
      cudaFree(f->H);
      delete f;
    }
  }
  delete e;
  return BBX_OK;
}

uint32_t bbx_engine_get_ring_length(const bbx_engine* e) { return e ? e->Rd : 0; }
void* bbx_engine_get_stream(const bbx_engine* e) { return e ? (void*)e->stream : nullptr; }

int bbx_filter_create(bbx_engine* e, const float* ir, uint32_t length, bbx_filter** out) {
  BBX_REQUIRE(e && out, "bbx_filter_create: null argument");
  BBX_REQUIRE(ir || length == 0, "bbx_filter_create: null impulse response");
  const uint32_t B = e->B, N = 2 * B;
  uint32_t P = std::max(1u, ceil_div(length, B));
  BBX_REQUIRE(P <= e->Pmax, "impulse response of %u taps needs %u partitions, engine max_partitions is %u", length, P, e->Pmax);
  DeviceGuard dg(e->device);
  // zero-padded windows [h[pB .. pB+B-1], 0^B]
  std::vector<float> pad((size_t)P * N, 0.0f);
  for (uint32_t i = 0; i < length; i++) pad[(size_t)(i / B) * N + (i % B)] = ir[i];
  float* d_pad = nullptr;
  bbx_filter* f = new bbx_filter();
  f->engine = e;
  f->P = P;
  f->H = nullptr;
  int rc = [&]() -> int {
    BBX_CUDA_TRY(cudaMalloc((void**)&d_pad, sizeof(float) * pad.size()));
    BBX_CUDA_TRY(cudaMalloc((void**)&f->H, sizeof(float2) * (size_t)P * B));
    BBX_CUDA_TRY(cudaMemcpyAsync(d_pad, pad.data(), sizeof(float) * pad.size(), cudaMemcpyHostToDevice, e->stream));
    // H = R2C(window) / N : the only normalisation of the whole path, exact (power of two)
    int lrc = launch_rfft(B, d_pad, 0, N, f->H, 0, P, 0, e->tw, 1.0f / (float)N, 1, P, e->stream);
    e->launches++;
    cudaError_t se = cudaStreamSynchronize(e->stream);
    if (!lrc && se != cudaSuccess) {
      set_error("filter transform failed: %s", cudaGetErrorString(se));
      lrc = BBX_ERR_CUDA;
    }
    return lrc;
  }();
  cudaFree(d_pad);
  if (rc) {
    cudaFree(f->H);
    delete f;
    return rc;
  }
  e->filters.push_back(f);
  {
    std::lock_guard<std::mutex> lk(g_filter_mu);
    g_live_filters.insert(f);
  }
  *out = f;
  return BBX_OK;
}

int bbx_filter_read_spectra(const bbx_filter* f, float* out, size_t max_floats) {
  BBX_REQUIRE(f && out, "bbx_filter_read_spectra: null argument");
  const size_t n = (size_t)f->P * f->engine->B * 2;
  BBX_REQUIRE(max_floats >= n, "bbx_filter_read_spectra: %zu floats needed, %zu given", n, max_floats);
  DeviceGuard dg(f->engine->device);
  BBX_CUDA_TRY(cudaMemcpy(out, f->H, sizeof(float) * n, cudaMemcpyDeviceToHost));
  return BBX_OK;
}

int bbx_filter_destroy(bbx_filter* f) {
  if (!f) return BBX_OK;
  {
    std::lock_guard<std::mutex> lk(g_filter_mu);
    if (!g_live_filters.count(f)) return BBX_OK;  // already released together with its engine
  }
  bbx_engine* e = f->engine;
  // The MAC plans and the tensor-core operand pack hold raw device pointers into the spectra of every selected filter:
  // a filter that a path still uses (current or latched) cannot go.  Select another filter (or NULL) first and run one
  // bbx_process call so that the switch has been applied, or destroy the engine, which releases all its filters.
  for (size_t k = 0; k < e->paths.size(); k++) {
    const PathState& p = e->paths[k];
    if (p.cur == f || (p.has_pending && p.pend == f)) {
      set_error("bbx_filter_destroy: the filter is still selected on path %zu (%s); select another filter and process a block first",
                k, p.cur == f ? "current" : "latched");
      return BBX_ERR_STATE;
    }
  }
  DeviceGuard dg(e->device);
  cudaStreamSynchronize(e->stream);
  e->filters.erase(std::remove(e->filters.begin(), e->filters.end(), f), e->filters.end());
  {
    std::lock_guard<std::mutex> lk(g_filter_mu);
    g_live_filters.erase(f);
  }
  cudaFree(f->H);
  delete f;
  return BBX_OK;
}

uint32_t bbx_filter_partitions(const bbx_filter* f) { return f ? f->P : 0; }

int bbx_set_route(bbx_engine* e, uint32_t path, uint32_t input, uint32_t output, float gain) {
  BBX_REQUIRE(e != nullptr, "bbx_set_route: null engine");
  BBX_REQUIRE(e->mode == BBX_MODE_ROUTED, "bbx_set_route: engine is not in ROUTED mode");
  BBX_REQUIRE(path < e->n_paths && input < e->n_in && output < e->n_out, "bbx_set_route: index out of range");
  PathState& p = e->paths[path];
  if (p.input != input) e->steady_dirty = true;
  p.input = input;
  p.output = output;
  p.gain = gain;
  e->route_dirty = true;
  return BBX_OK;
}

int bbx_set_filter(bbx_engine* e, uint32_t path, const bbx_filter* filter, int crossfade, double delay) {
  BBX_REQUIRE(e != nullptr, "bbx_set_filter: null engine");
  BBX_REQUIRE(path < e->n_paths, "bbx_set_filter: path %u out of range", path);
  BBX_REQUIRE(!filter || filter->engine == e, "bbx_set_filter: filter belongs to another engine");
  BBX_REQUIRE(delay >= 0.0 && delay <= (double)e->cfg.max_delay, "bbx_set_filter: delay %g outside [0, max_delay=%u]", delay,
              e->cfg.max_delay);
  BBX_REQUIRE(e->mode != BBX_MODE_MIMO || delay == 0.0, "bbx_set_filter: MIMO mode has no per-path delay");
  BBX_REQUIRE(!(e->sh_world > 1 || e->comm) || !crossfade,
              "bbx_set_filter: the input-sharded MIMO engine switches filters without crossfade");
  PathState& p = e->paths[path];
  p.pend = filter;
  p.pend_delay = delay;
  p.xfade = crossfade != 0;
  p.has_pending = true;
  return BBX_OK;
}

int bbx_set_filters(bbx_engine* e, uint32_t n, const uint32_t* paths, const bbx_filter* const* filters, const int* crossfade,
                    const double* delays) {
  BBX_REQUIRE(e != nullptr && (n == 0 || (paths && filters)), "bbx_set_filters: null argument");
  // validate everything first: either all n switches are latched or none
  for (uint32_t k = 0; k < n; k++) {
    const double d = delays ? delays[k] : 0.0;
    BBX_REQUIRE(paths[k] < e->n_paths, "bbx_set_filters: path %u out of range", paths[k]);
    BBX_REQUIRE(!filters[k] || filters[k]->engine == e, "bbx_set_filters: filter %u belongs to another engine", k);
    BBX_REQUIRE(d >= 0.0 && d <= (double)e->cfg.max_delay, "bbx_set_filters: delay %g outside [0, max_delay=%u]", d, e->cfg.max_delay);
    BBX_REQUIRE(e->mode != BBX_MODE_MIMO || d == 0.0, "bbx_set_filters: MIMO mode has no per-path delay");
    BBX_REQUIRE(!(e->sh_world > 1 || e->comm) || !(crossfade && crossfade[k]),
                "bbx_set_filters: the input-sharded MIMO engine switches filters without crossfade");
  }
  for (uint32_t k = 0; k < n; k++) {
    PathState& p = e->paths[paths[k]];
    p.pend = filters[k];
    p.pend_delay = delays ? delays[k] : 0.0;
    p.xfade = crossfade && crossfade[k] != 0;
    p.has_pending = true;
  }
  return BBX_OK;
}

// launch KERNEL<FMT, ACC> for a runtime (format, big-endian, typed-access) triple; ACC as in formats.cuh
#define BBX_PCM_LAUNCH_ACC(KERNEL, FMT, be, fast, GRID, STREAM, ARGS)            \
  do {                                                                            \
    if ((fast) && FMT != FMT_24) KERNEL<FMT, 2><<<GRID, 256, 0, STREAM>>>(ARGS);  \
    else if (be) KERNEL<FMT, 1><<<GRID, 256, 0, STREAM>>>(ARGS);                  \
    else KERNEL<FMT, 0><<<GRID, 256, 0, STREAM>>>(ARGS);                          \
  } while (0)
#define BBX_PCM_LAUNCH(KERNEL, fmt, be, fast, GRID, STREAM, ARGS)                          \
  do {                                                                                      \
    switch (fmt) {                                                                          \
      case FMT_16: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_16, be, fast, GRID, STREAM, ARGS); break;   \
      case FMT_24: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_24, be, fast, GRID, STREAM, ARGS); break;   \
      case FMT_32: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_32, be, fast, GRID, STREAM, ARGS); break;   \
      case FMT_F32: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_F32, be, fast, GRID, STREAM, ARGS); break; \
      default: BBX_PCM_LAUNCH_ACC(KERNEL, FMT_F64, be, fast, GRID, STREAM, ARGS); break;      \
    }                                                                                       \
  } while (0)

static int launch_pcm_in(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, uint32_t T, cudaStream_t st) {
  const uint32_t B = e->B;
  {
    PcmInArgs a;    a.pcm = (const uint8_t*)in;
    a.fmt = infmt;
    a.be = in_be;
    a.in_channels = in_channels;
    a.n_inputs = e->n_in;
    a.B = B;
    a.T = T;
    a.xin_cur = e->xin[e->parity];
    a.xin_prev = e->xin[e->parity ^ 1];
    a.xstride = e->xstride;
    a.prev_off = e->tprev * B;
    {
      const uint32_t bps = fmt_bytes(infmt);
      a.fast = (!in_be && bps != 3 && ((uintptr_t)in % bps) == 0) ? 1 : 0;  // frame stride = in_channels * bps is aligned too
    }
    // wide tiles pay when the channel axis fills the lanes; few-channel engines keep the finer grid
    if (B % 128 == 0 && e->n_in >= 16)
      BBX_PCM_LAUNCH(k_pcm_in128, infmt, in_be, a.fast, dim3((T + 1) * B / 128, ceil_div(e->n_in, 32)), st, a);
    else BBX_PCM_LAUNCH(k_pcm_in, infmt, in_be, a.fast, dim3((T + 1) * B / 32, ceil_div(e->n_in, 32)), st, a);
    BBX_CUDA_TRY(cudaGetLastError());
    e->launches++;
  }
  return BBX_OK;
}

// argument and geometry checks shared by every form of bbx_process: they run before anything is staged, copied or latched
static int validate_call(const bbx_engine* e, const void* in, int infmt, uint32_t in_channels, const void* out, int outfmt,
                         uint32_t out_channels, uint32_t nframes) {
  BBX_REQUIRE(e && in && out, "bbx_process: null argument");
  BBX_REQUIRE(infmt > FMT_UNKNOWN && infmt < FMT_COUNT && outfmt > FMT_UNKNOWN && outfmt < FMT_COUNT, "bbx_process: bad format");
  BBX_REQUIRE(nframes > 0 && nframes % e->B == 0, "bbx_process: nframes %u is not a positive multiple of the block size %u",
              nframes, e->B);
  BBX_REQUIRE(nframes / e->B <= e->Tmax, "bbx_process: %u blocks exceed max_blocks %u", nframes / e->B, e->Tmax);
  BBX_REQUIRE(in_channels >= e->n_in && out_channels >= e->n_out_pcm, "bbx_process: too few channels in the PCM buffers");
  BBX_REQUIRE(e->sh_world <= 1 || e->comm || e->px_on,
              "bbx_process: the input-sharded MIMO engine needs bbx_engine_set_comm() or bbx_engine_peer_attach()");
  return BBX_OK;
}

int bbx_process_dev(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt,
                    int out_be, uint32_t out_channels, uint32_t nframes) {
  int rc = validate_call(e, in, infmt, in_channels, out, outfmt, out_channels, nframes);
  if (rc) return rc;
  NvtxRange nv_call("bbx_process_dev");
  const uint32_t B = e->B, T = nframes / B;
  DeviceGuard dg(e->device);
  cudaStream_t st = e->stream;
  e->last_infmt = infmt;
  e->last_outfmt = outfmt;

  // ---- latch pending switches: they apply at this call's first block boundary ----
  bool any_xfade = false;
  for (auto& p : e->paths)
    if (p.has_pending) {
      if (p.xfade) any_xfade = true;
      else {  // hard switch: filter and delay jump now
        if (p.cur != p.pend) e->steady_dirty = true;
        p.cur = p.pend;
        p.delay = p.pend_delay;
        p.has_pending = false;
        e->route_dirty = true;
      }
    }
  uint32_t n_first = 0;
  std::vector<std::vector<JobTerm>> jobs;
  if (any_xfade) {
    // transitional plan for block 0: old filters as the main jobs, new filters as extra jobs
    make_jobs(e, false, jobs);
    std::vector<uint32_t> xjob(e->n_streams, kNoJob);
    if (e->mode == BBX_MODE_MIMO) {
      for (uint32_t o = 0; o < e->n_out; o++) {
        bool sw = false, differs = false;
        std::vector<JobTerm> nj;
        for (uint32_t i = 0; i < e->n_in; i++) {
          const PathState& p = e->paths[(size_t)o * e->n_in + i];
          const bbx_filter* f = p.has_pending ? p.pend : p.cur;
          if (p.has_pending) {
            sw = true;
            if (p.pend != p.cur) differs = true;
          }
          if (f) nj.push_back({f, i});
        }
        if (sw) {
          if (differs) {
            xjob[o] = (uint32_t)jobs.size();
            jobs.push_back(nj);
          } else {
            xjob[o] = kSameJob;
          }
        }
      }
    } else {
      for (uint32_t k = 0; k < e->n_paths; k++) {
        const PathState& p = e->paths[k];
        if (!p.has_pending) continue;
        if (p.pend == p.cur) xjob[k] = kSameJob;
        else {
          xjob[k] = (uint32_t)jobs.size();
          std::vector<JobTerm> nj;
          if (p.pend) nj.push_back({p.pend, p.input});
          jobs.push_back(nj);
        }
      }
    }
    if ((rc = build_plan(e, e->plan_first, jobs, xjob))) return rc;
    n_first = 1;
    if ((rc = upload_routes(e, true))) return rc;
    e->route_dirty = true;  // the steady-state tables follow after this call
    // commit the crossfaded switches
    for (auto& p : e->paths)
      if (p.has_pending) {
        if (p.cur != p.pend) e->steady_dirty = true;
        p.cur = p.pend;
        p.delay = p.pend_delay;
        p.has_pending = false;
      }
  } else if (e->route_dirty) {
    if ((rc = upload_routes(e, false))) return rc;
    e->route_dirty = false;
  }
  if (e->steady_dirty || !e->plan_steady.valid) {
    make_jobs(e, false, jobs);
    rc = build_plan(e, e->plan_steady, jobs, std::vector<uint32_t>());
    if (rc) return rc;
    e->steady_dirty = false;
    e->tc_dirty = true;
  }
  // MIMO calls of >= tc_min_blocks blocks without a crossfade in flight run the per-bin GEMM on the tensor cores
  const bool use_tc = e->tc_on && n_first == 0 && T >= e->tc_min_blocks && e->plan_steady.total_rows > 0;
  if (use_tc && e->tc_dirty) {
    if ((rc = tc_pack_filters(e))) return rc;
  }

  // ---- streaming call of a PER_CHANNEL / ROUTED engine with short filters: one launch does steps 1 - 4 (and 5) ----
  const MacPlan& fpl = n_first ? e->plan_first : e->plan_steady;
  const bool fuse = e->fused_on && T == 1 && e->mode != BBX_MODE_MIMO && e->sh_world <= 1 && !e->comm && fpl.max_job_rows <= e->fused_max_rows;
  const uint32_t ibps = fmt_bytes(infmt), obps = fmt_bytes(outfmt);
  // PCM in host memory behind PCIe: a stream's strided picks out of wide frames would be one small bus request each, so
  // wide host layouts go through the transposing PCM kernels (lanes over channels) on that side
  const bool planar_in = fuse && e->in_is_host && (size_t)in_channels * ibps > 32;
  const bool fuse_out = fuse && e->mode == BBX_MODE_PER_CHANNEL && !(e->out_is_host && (size_t)out_channels * obps > 32);
  if (fuse) {
    NvtxRange nv("bbx.block_fused");
    FusedArgs a;
    if (planar_in) {
      if ((rc = launch_pcm_in(e, in, infmt, in_be, in_channels, T, st))) return rc;
    }
    a.planar_in = planar_in ? 1 : 0;
    a.pcm_in = (const uint8_t*)in;
    a.pcm_out = (uint8_t*)out;
    a.infmt = infmt;
    a.in_be = in_be;
    a.in_fast = (!in_be && ibps != 3 && ((uintptr_t)in % ibps) == 0) ? 1 : 0;
    a.outfmt = outfmt;
    a.out_be = out_be;
    a.out_fast = (!out_be && obps != 3 && ((uintptr_t)out % obps) == 0) ? 1 : 0;
    a.in_channels = in_channels;
    a.out_channels = out_channels;
    a.n_streams = e->n_streams;
    a.stream_input = (e->mode == BBX_MODE_ROUTED) ? (const uint32_t*)(e->d_route + e->roff_input) : nullptr;
    a.xin_cur = e->xin[e->parity];
    a.xin_prev = e->xin[e->parity ^ 1];
    a.xstride = e->xstride;
    a.prev_off = e->tprev * B;
    a.fdl = e->fdl;
    a.R = e->R;
    a.head = e->head;
    a.tw = e->tw;
    a.segs = fpl.segs();
    a.job_seg_first = fpl.job_seg_first();
    a.xjob = n_first ? fpl.view().xjob : nullptr;
    a.ybuf = e->ybuf;
    a.Rd = e->Rd;
    a.wpos = e->wpos;
    a.fractional = e->cfg.fractional_delay;
    a.entry = route_view(e).entry;
    if ((rc = fuse_out ? launch_fused<true>(B, a, st) : launch_fused<false>(B, a, st))) return rc;
    e->launches++;
    e->fused_calls++;
    e->last_mac_kernel = "k_block_fused";
  } else {
  // ---- 1. PCM -> planar fp32 ----
  {
    NvtxRange nv("bbx.pcm_in");
    if ((rc = launch_pcm_in(e, in, infmt, in_be, in_channels, T, st))) return rc;
  }
  // ---- 2. forward transforms into the FDL ----
  {
    NvtxRange nv("bbx.rfft");
    if ((rc = launch_rfft(B, e->xin[e->parity], e->xstride, B, e->fdl, (uint64_t)e->R * B, e->R, e->head, e->tw, 1.0f, e->n_in, T, st)))
      return rc;
    e->launches++;
  }
  // ---- 3. FDL multiply-accumulate ----
  {
    NvtxRange nv(use_tc ? "bbx.mac_tensor" : "bbx.mac");
    if (use_tc) {
      if ((rc = launch_mimo_tc(e, T))) return rc;
    } else if (n_first) {
      if ((rc = launch_mac(e, e->plan_first, 0, 1))) return rc;
      if ((rc = launch_mac(e, e->plan_steady, 1, T - 1))) return rc;
    } else {
      if ((rc = launch_mac(e, e->plan_steady, 0, T))) return rc;
    }
  }
  // ---- 3b. input-sharded MIMO: sum the partial spectra over the ranks, keep the local outputs ----
  if (e->sh_world > 1 || e->comm) {
    NvtxRange nv("bbx.exchange");
    const PlanView pv = use_tc ? tc_plan_view(e) : e->plan_steady.view();
    cudaEvent_t x0 = nullptr, x1 = nullptr;
    if (e->profile_mac) {
      if (e->xchg_events_used + 2 > e->xchg_events.size()) {
        size_t old = e->xchg_events.size();
        e->xchg_events.resize(old + 64);
        for (size_t i = old; i < e->xchg_events.size(); i++) BBX_CUDA_TRY(cudaEventCreate(&e->xchg_events[i]));
      }
      x0 = e->xchg_events[e->xchg_events_used++];
      x1 = e->xchg_events[e->xchg_events_used++];
      BBX_CUDA_TRY(cudaEventRecord(x0, st));
      e->xchg_count++;
      // bytes this rank sends to its peers: the partial spectra of every output it does not own
      e->xchg_bytes += (uint64_t)(e->n_out - e->sh_nloc) * T * B * sizeof(float2);
    }
    if (e->px_on) {
      e->px_epoch++;
      const uint32_t parity = e->px_epoch & 1u;
      k_gather_spectra_peer<<<dim3(e->n_out, ceil_div(T, kPeerTPB)), 256, 0, st>>>(e->ypart, use_tc ? nullptr : e->nyq_part, e->max_slots, pv,
                                                               e->px_table, e->sh_world, e->sh_rank, e->sh_nloc, B, T, e->px_half,
                                                               parity, e->px_epoch, e->px_done);
      BBX_CUDA_TRY(cudaGetLastError());
      k_peer_wait<<<1, 32, 0, st>>>((const uint32_t*)(e->px_mem + e->px_flag_off), e->sh_world, parity, e->px_epoch, e->px_status);
      BBX_CUDA_TRY(cudaGetLastError());
      e->launches += 2;
    } else {
      k_gather_spectra<<<dim3(e->n_out, T), 256, 0, st>>>(e->ypart, use_tc ? nullptr : e->nyq_part, e->max_slots, pv, e->sh_send, B, T);
      BBX_CUDA_TRY(cudaGetLastError());
      e->launches++;
      if ((rc = comm_reduce_scatter_f32(e->comm, (const float*)e->sh_send, (float*)e->sh_recv, (size_t)e->sh_nloc * T * B * 2, st)))
        return rc;
    }
    if (x1) BBX_CUDA_TRY(cudaEventRecord(x1, st));
  }
  // ---- 4. inverse transforms, crossfade, delay ring ----
  {
    NvtxRange nv("bbx.irfft");
    if ((rc = launch_irfft(e, T, n_first, use_tc, st))) return rc;
    e->launches++;
  }
  }  // !fuse
  if (!fuse_out) {
  // ---- 5. delay read, mixdown, output format ----
  {
    NvtxRange nv("bbx.pcm_out");
    PcmOutArgs a;
    const uint32_t obps = fmt_bytes(outfmt);
    a.pcm = (uint8_t*)out;
    a.fmt = outfmt;
    a.be = out_be;
    a.out_channels = out_channels;
    a.n_outputs = e->n_out_pcm;
    a.B = B;
    a.T = T;
    a.ybuf = e->ybuf;
    a.Rd = e->Rd;
    a.wpos0 = e->wpos;
    a.fractional = e->cfg.fractional_delay;
    a.fast = (!out_be && obps != 3 && ((uintptr_t)out % obps) == 0) ? 1 : 0;
    a.rv = route_view(e);
    // wide tiles for many outputs with integer delays (their ring reads batch); mixdowns of many paths into few
    // outputs and fractional delays (14-tap double-precision reads) keep the finer grid
    // mixdowns (at least four paths per output on average, few outputs): the (route, frame)-parallel kernel
    const bool mix = e->n_out_pcm <= 32 && e->n_routes_pcm <= kMixMaxRoutes && e->n_routes_pcm >= 4 * e->n_out_pcm &&
                     e->pcm_out_mix;
    if (mix) BBX_PCM_LAUNCH(k_pcm_out_mix, outfmt, out_be, a.fast, dim3(T * B / 32), st, a);
    else if (e->cfg.fractional_delay && e->pcm_out_mix)  // one sample per thread (the same switch keeps it on the tile kernel)
      BBX_PCM_LAUNCH(k_pcm_out_frac, outfmt, out_be, a.fast, dim3(T * B / 8, ceil_div(e->n_out_pcm, 32)), st, a);
    else if (B % 128 == 0 && e->n_out_pcm >= 16 && !e->cfg.fractional_delay)
      BBX_PCM_LAUNCH(k_pcm_out128, outfmt, out_be, a.fast, dim3(T * B / 128, ceil_div(e->n_out_pcm, 32)), st, a);
    else BBX_PCM_LAUNCH(k_pcm_out, outfmt, out_be, a.fast, dim3(T * B / 32, ceil_div(e->n_out_pcm, 32)), st, a);
    BBX_CUDA_TRY(cudaGetLastError());
    e->launches++;
  }
  }  // !fuse_out
  // ---- advance the state ----
  e->head = (e->head + T) % e->R;
  e->wpos = (e->wpos + T * B) % e->Rd;
  e->parity ^= 1;
  e->tprev = T;
  return BBX_OK;
}

// Device-side address of a pinned host buffer (cudaHostAlloc / cudaHostRegister under unified addressing), or false.
static bool mapped_device_ptr(const void* host, void** dev) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
  *dev = at.devicePointer;
  return true;
}

int bbx_process_async(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt,
                      int out_be, uint32_t out_channels, uint32_t nframes) {
  {
    const int vrc = validate_call(e, in, infmt, in_channels, out, outfmt, out_channels, nframes);
    if (vrc) return vrc;  // nothing staged, no copy queued
  }
  NvtxRange nv_call("bbx_process_async");
  DeviceGuard dg(e->device);
  const uint32_t ibps = fmt_bytes(infmt), obps = fmt_bytes(outfmt);
  size_t in_bytes = (size_t)nframes * in_channels * ibps;
  size_t out_bytes = (size_t)nframes * out_channels * obps;
  size_t need = std::max(in_bytes, out_bytes);
  // Latency path (real-time callers: one or a few blocks per call): a short buffer in pinned host memory that the
  // device can address is read by k_pcm_in / written by k_pcm_out straight over PCIe -- no copy-engine operation and no
  // cross-stream hand-off on that side.  Decided per side, and only where the kernel's accesses suit the bus: typed
  // little-endian samples (no 3-byte formats) and at least 128 contiguous bytes of used channels per frame (a warp
  // covers 32 channels of one frame); narrow or byte-wise layouts would turn into many small PCIe transactions and
  // stay on the staged path, like pageable buffers.
  void *din = nullptr, *dout = nullptr;
  const bool typed_in = !in_be && ibps != 3 && ((uintptr_t)in % ibps) == 0;
  const bool typed_out = !out_be && obps != 3 && ((uintptr_t)out % obps) == 0;
  // a side suits the bus when a warp's accesses fill whole PCIe requests: at least 128 bytes of used channels per frame, or
  // frames of at most 32 bytes altogether (stereo float: consecutive frames share the sectors, nothing is fetched in vain)
  const bool dense_in = (size_t)e->n_in * ibps >= 128 || (size_t)in_channels * ibps <= 32;
  const bool dense_out = (size_t)e->n_out_pcm * obps >= 128 || (size_t)out_channels * obps <= 32;
  const bool direct_in = in_bytes <= e->direct_io_max_bytes && typed_in && dense_in && mapped_device_ptr(in, &din);
  const bool direct_out = out_bytes <= e->direct_io_max_bytes && typed_out && dense_out && mapped_device_ptr(out, &dout);
  if (direct_in && direct_out) {
    e->direct_calls++;
    e->in_is_host = e->out_is_host = true;
    const int drc = bbx_process_dev(e, din, infmt, in_be, in_channels, dout, outfmt, out_be, out_channels, nframes);
    e->in_is_host = e->out_is_host = false;
    return drc;
  }
  if (need <= e->direct_io_max_bytes) {
    // Latency mode for the sides that are staged after all (24-bit or sparse layouts, pageable memory): the copies go on
    // the engine stream itself into staging buffers of their own -- no copy streams, no cross-stream events; a call this
    // small has nothing to overlap, and the event hops cost more than its kernels.
    if (e->d_lat_bytes < need) {
      BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
      cudaFree(e->d_lat_in);
      cudaFree(e->d_lat_out);
      e->d_lat_in = e->d_lat_out = nullptr;
      e->d_lat_bytes = std::max<size_t>(need, 64 << 10);
      BBX_CUDA_TRY(cudaMalloc((void**)&e->d_lat_in, e->d_lat_bytes));
      BBX_CUDA_TRY(cudaMalloc((void**)&e->d_lat_out, e->d_lat_bytes));
    }
    if (direct_in || direct_out) e->direct_calls++;
    e->in_is_host = direct_in;
    e->out_is_host = direct_out;
    if (!direct_in) BBX_CUDA_TRY(cudaMemcpyAsync(e->d_lat_in, in, in_bytes, cudaMemcpyHostToDevice, e->stream));
    if (!direct_out && out_channels > e->n_out_pcm)  // channels beyond n_outputs keep the caller's bytes
      BBX_CUDA_TRY(cudaMemcpyAsync(e->d_lat_out, out, out_bytes, cudaMemcpyHostToDevice, e->stream));
    int lrc = bbx_process_dev(e, direct_in ? din : e->d_lat_in, infmt, in_be, in_channels, direct_out ? dout : e->d_lat_out, outfmt,
                              out_be, out_channels, nframes);
    e->in_is_host = e->out_is_host = false;
    if (lrc) return lrc;
    if (!direct_out) BBX_CUDA_TRY(cudaMemcpyAsync(out, e->d_lat_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
    return BBX_OK;
  }
  if (!e->d_in[0] || e->d_io_bytes < need) {
    BBX_CUDA_TRY(cudaStreamSynchronize(e->s_in));
    BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
    BBX_CUDA_TRY(cudaStreamSynchronize(e->s_out));
    e->d_io_bytes = std::max(need, e->d_io_bytes);
    for (int i = 0; i < bbx_engine::kIoSlots; i++) {
      cudaFree(e->d_in[i]);
      cudaFree(e->d_out[i]);
      e->d_in[i] = e->d_out[i] = nullptr;
      BBX_CUDA_TRY(cudaMalloc((void**)&e->d_in[i], e->d_io_bytes));
      BBX_CUDA_TRY(cudaMalloc((void**)&e->d_out[i], e->d_io_bytes));
    }
  }
  const int k = (int)(e->host_calls % bbx_engine::kIoSlots);
  e->host_calls++;
  e->copy_streams_busy = true;
  if (direct_in || direct_out) e->direct_calls++;
  // H2D on the input-copy stream, once the kernels of call n - kIoSlots have finished reading this staging buffer
  bool fed = false;
  cudaEvent_t* tr = (e->io_calls < e->io_cap) ? &e->io_ev[6 * e->io_calls++] : nullptr;
  if (!direct_in) {
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_in, e->ev_comp[k], 0));
    if (tr) BBX_CUDA_TRY(cudaEventRecord(tr[0], e->s_in));
    BBX_CUDA_TRY(cudaMemcpyAsync(e->d_in[k], in, in_bytes, cudaMemcpyHostToDevice, e->s_in));
    if (tr) BBX_CUDA_TRY(cudaEventRecord(tr[1], e->s_in));
    fed = true;
  }
  if (!direct_out && out_channels > e->n_out_pcm) {
    // channels beyond n_outputs keep the caller's bytes: seed the output staging with them
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_in, e->ev_d2h[k], 0));
    BBX_CUDA_TRY(cudaMemcpyAsync(e->d_out[k], out, out_bytes, cudaMemcpyHostToDevice, e->s_in));
    fed = true;
  }
  if (fed) {
    BBX_CUDA_TRY(cudaEventRecord(e->ev_h2d[k], e->s_in));
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_h2d[k], 0));
  }
  // kernels on the engine stream
  if (!direct_out) BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_d2h[k], 0));
  if (tr) BBX_CUDA_TRY(cudaEventRecord(tr[2], e->stream));
  int rc = bbx_process_dev(e, direct_in ? din : e->d_in[k], infmt, in_be, in_channels, direct_out ? dout : e->d_out[k], outfmt, out_be,
                           out_channels, nframes);
  if (rc) return rc;
  if (tr) BBX_CUDA_TRY(cudaEventRecord(tr[3], e->stream));
  BBX_CUDA_TRY(cudaEventRecord(e->ev_comp[k], e->stream));
  if (!direct_out) {
    // D2H on the output-copy stream
    BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_out, e->ev_comp[k], 0));
    if (tr) BBX_CUDA_TRY(cudaEventRecord(tr[4], e->s_out));
    BBX_CUDA_TRY(cudaMemcpyAsync(out, e->d_out[k], out_bytes, cudaMemcpyDeviceToHost, e->s_out));
    if (tr) BBX_CUDA_TRY(cudaEventRecord(tr[5], e->s_out));
    BBX_CUDA_TRY(cudaEventRecord(e->ev_d2h[k], e->s_out));
  }
  return BBX_OK;
}

int bbx_engine_sync(bbx_engine* e) {
  BBX_REQUIRE(e != nullptr, "bbx_engine_sync: null engine");
  DeviceGuard dg(e->device);
  // the copy streams only carry work of the staged host path and of the timer: a latency call (everything on the engine
  // stream) does not pay two driver calls for idle streams
  const bool copies = e->copy_streams_busy;
  if (copies) BBX_CUDA_TRY(cudaStreamSynchronize(e->s_in));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (copies) {
    BBX_CUDA_TRY(cudaStreamSynchronize(e->s_out));
    e->copy_streams_busy = false;
  }
  if (e->px_status_h && *(volatile int*)e->px_status_h) {
    set_error("peer mixdown: rank %d never published its partial spectra (timed out); results are invalid",
              *(volatile int*)e->px_status_h - 1);
    return BBX_ERR_CUDA;
  }
  if (e->mac_status_h && *(volatile int*)e->mac_status_h) {
    set_error("k_fdl_mac_tbs: a barrier wait timed out inside the time-batched MAC (status %d); results are invalid",
              *(volatile int*)e->mac_status_h);
    return BBX_ERR_CUDA;
  }
  if (e->tc_status_h && *(volatile int*)e->tc_status_h) {
    // fail loudly: the output of that call is not valid
    set_error("k_mimo_tc: a barrier wait timed out inside the tensor-core kernel (status %d); results are invalid",
              *(volatile int*)e->tc_status_h);
    return BBX_ERR_CUDA;
  }
  return BBX_OK;
}

int bbx_process(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt,
                int out_be, uint32_t out_channels, uint32_t nframes) {
  int rc = bbx_process_async(e, in, infmt, in_be, in_channels, out, outfmt, out_be, out_channels, nframes);
  if (rc) return rc;
  return bbx_engine_sync(e);
}

int bbx_block_latency(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt, int out_be,
                      uint32_t out_channels, uint32_t nframes, uint32_t ncalls, uint32_t warmup, double* us) {
  BBX_REQUIRE(e && us && ncalls, "bbx_block_latency: null argument");
  for (uint32_t i = 0; i < warmup + ncalls; i++) {
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const int rc = bbx_process(e, in, infmt, in_be, in_channels, out, outfmt, out_be, out_channels, nframes);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (rc) return rc;
    if (i >= warmup) us[i - warmup] = (double)(t1.tv_sec - t0.tv_sec) * 1e6 + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-3;
  }
  return BBX_OK;
}

int bbx_blockconvolver_convolve(bbx_engine* e, const float* in, float* out) {
  BBX_REQUIRE(e && in && out, "bbx_blockconvolver_convolve: null argument");
  BBX_REQUIRE(e->n_in == 1 && e->n_out == 1, "bbx_blockconvolver_convolve: engine must be single-channel");
  return bbx_process(e, in, FMT_F32, 0, 1, out, FMT_F32, 0, 1, e->B);
}

int bbx_engine_timer_start(bbx_engine* e) {
  BBX_REQUIRE(e != nullptr, "null engine");
  // nothing of the timed region may start before the start event: drain, record, and fence the other streams
  int rc = bbx_engine_sync(e);
  if (rc) return rc;
  e->copy_streams_busy = true;
  BBX_CUDA_TRY(cudaEventRecord(e->ev_start, e->s_in));
  BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_start, 0));
  BBX_CUDA_TRY(cudaStreamWaitEvent(e->s_out, e->ev_start, 0));
  return BBX_OK;
}
int bbx_engine_timer_stop(bbx_engine* e, float* elapsed_ms) {
  BBX_REQUIRE(e && elapsed_ms, "null argument");
  // the stop event follows everything enqueued on the copy streams and the engine stream
  BBX_CUDA_TRY(cudaEventRecord(e->ev_join_in, e->s_in));
  BBX_CUDA_TRY(cudaEventRecord(e->ev_join_out, e->s_out));
  BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join_in, 0));
  BBX_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join_out, 0));
  BBX_CUDA_TRY(cudaEventRecord(e->ev_stop, e->stream));
  BBX_CUDA_TRY(cudaEventSynchronize(e->ev_stop));
  BBX_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, e->ev_start, e->ev_stop));
  return BBX_OK;
}
uint64_t bbx_engine_launch_count(const bbx_engine* e) { return e ? e->launches : 0; }
const char* bbx_engine_mac_kernel(const bbx_engine* e) { return e ? e->last_mac_kernel : "none"; }

int bbx_engine_profile_mac(bbx_engine* e, int enable) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->profile_mac = enable != 0;
  e->mac_events_used = 0;
  e->mac_ms_total = 0.0;
  e->mac_launches = e->mac_units = e->mac_bytes = 0;
  e->xchg_events_used = 0;
  e->xchg_ms_total = 0.0;
  e->xchg_count = e->xchg_bytes = 0;
  return BBX_OK;
}

// ---- checkpoint / resume of an engine's audio state --------------------------------------------------------------------
// Everything a later call depends on: the FDL ring, the previous call's last input block, the delay rings, the ring
// positions, and per path the selected / latched filter, delays and gain.  Filters are referred to by their position in
// the engine's filter list (creation order among the live ones): the restoring engine must hold the same filters in the
// same order.  Plans, routes and the tensor-core operand pack are derived data and are rebuilt by the next call.
namespace {
struct StateHeader {
  uint32_t magic, version;
  uint32_t B, Pmax, n_in, n_out, n_paths, n_streams, Tmax, R, Rd, xstride, mode, n_filters;
  uint32_t head, wpos, parity, tprev;
  uint64_t xin_bytes, fdl_bytes, ybuf_bytes, path_bytes;
};
struct StatePath {
  int32_t cur, pend;  // index into the filter list, -1 = none
  uint32_t has_pending, xfade;
  float gain;
  uint32_t pad;
  double delay, pend_delay;
};
constexpr uint32_t kStateMagic = 0x58424253u /* "SBBX" */, kStateVersion = 1;

StateHeader state_header(const bbx_engine* e) {
  StateHeader h;
  memset(&h, 0, sizeof(h));
  h.magic = kStateMagic, h.version = kStateVersion;
  h.B = e->B, h.Pmax = e->Pmax, h.n_in = e->n_in, h.n_out = e->n_out, h.n_paths = e->n_paths, h.n_streams = e->n_streams;
  h.Tmax = e->Tmax, h.R = e->R, h.Rd = e->Rd, h.xstride = e->xstride, h.mode = (uint32_t)e->mode;
  h.n_filters = (uint32_t)e->filters.size();
  h.head = e->head, h.wpos = e->wpos, h.parity = e->parity, h.tprev = e->tprev;
  h.xin_bytes = sizeof(float) * (uint64_t)e->n_in * e->xstride;
  h.fdl_bytes = sizeof(float2) * (uint64_t)e->n_in * e->R * e->B;
  h.ybuf_bytes = sizeof(float) * (uint64_t)e->n_streams * e->Rd;
  h.path_bytes = sizeof(StatePath) * (uint64_t)e->paths.size();
  return h;
}
size_t state_bytes(const StateHeader& h) { return sizeof(StateHeader) + h.path_bytes + 2 * h.xin_bytes + h.fdl_bytes + h.ybuf_bytes; }
int filter_index(const bbx_engine* e, const bbx_filter* f) {
  if (!f) return -1;
  for (size_t i = 0; i < e->filters.size(); i++)
    if (e->filters[i] == f) return (int)i;
  return -1;
}
int drain(bbx_engine* e) {
  BBX_CUDA_TRY(cudaStreamSynchronize(e->s_in));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->s_out));
  return BBX_OK;
}
}  // namespace

int bbx_engine_state_size(const bbx_engine* e, size_t* bytes) {
  BBX_REQUIRE(e && bytes, "bbx_engine_state_size: null argument");
  *bytes = state_bytes(state_header(e));
  return BBX_OK;
}

int bbx_engine_get_state(bbx_engine* e, void* buf, size_t capacity) {
  BBX_REQUIRE(e && buf, "bbx_engine_get_state: null argument");
  BBX_REQUIRE(e->sh_world <= 1 && !e->comm && !e->px_on, "bbx_engine_get_state: not available for an input-sharded engine");
  const StateHeader h = state_header(e);
  BBX_REQUIRE(capacity >= state_bytes(h), "bbx_engine_get_state: buffer of %zu bytes, state needs %zu", capacity, state_bytes(h));
  DeviceGuard dg(e->device);
  int rc = drain(e);
  if (rc) return rc;
  uint8_t* p = (uint8_t*)buf;
  memcpy(p, &h, sizeof(h));
  p += sizeof(h);
  for (const PathState& ps : e->paths) {
    StatePath sp;
    memset(&sp, 0, sizeof(sp));
    sp.cur = filter_index(e, ps.cur), sp.pend = filter_index(e, ps.pend);
    sp.has_pending = ps.has_pending, sp.xfade = ps.xfade;
    sp.gain = ps.gain, sp.delay = ps.delay, sp.pend_delay = ps.pend_delay;
    memcpy(p, &sp, sizeof(sp));
    p += sizeof(sp);
  }
  for (int i = 0; i < 2; i++, p += h.xin_bytes) BBX_CUDA_TRY(cudaMemcpy(p, e->xin[i], h.xin_bytes, cudaMemcpyDeviceToHost));
  BBX_CUDA_TRY(cudaMemcpy(p, e->fdl, h.fdl_bytes, cudaMemcpyDeviceToHost));
  p += h.fdl_bytes;
  BBX_CUDA_TRY(cudaMemcpy(p, e->ybuf, h.ybuf_bytes, cudaMemcpyDeviceToHost));
  return BBX_OK;
}

int bbx_engine_set_state(bbx_engine* e, const void* buf, size_t bytes) {
  BBX_REQUIRE(e && buf, "bbx_engine_set_state: null argument");
  BBX_REQUIRE(e->sh_world <= 1 && !e->comm && !e->px_on, "bbx_engine_set_state: not available for an input-sharded engine");
  BBX_REQUIRE(bytes >= sizeof(StateHeader), "bbx_engine_set_state: %zu bytes is not a state", bytes);
  StateHeader h;
  memcpy(&h, buf, sizeof(h));
  BBX_REQUIRE(h.magic == kStateMagic && h.version == kStateVersion, "bbx_engine_set_state: not an engine state (magic %08x, version %u)",
              h.magic, h.version);
  const StateHeader own = state_header(e);
  BBX_REQUIRE(h.B == own.B && h.Pmax == own.Pmax && h.n_in == own.n_in && h.n_out == own.n_out && h.n_paths == own.n_paths &&
                  h.n_streams == own.n_streams && h.Tmax == own.Tmax && h.R == own.R && h.Rd == own.Rd && h.xstride == own.xstride &&
                  h.mode == own.mode && h.path_bytes == own.path_bytes,
              "bbx_engine_set_state: the state was saved by an engine of a different geometry");
  BBX_REQUIRE(h.n_filters == own.n_filters, "bbx_engine_set_state: the state refers to %u filters, this engine holds %u", h.n_filters,
              own.n_filters);
  BBX_REQUIRE(bytes >= state_bytes(h), "bbx_engine_set_state: %zu bytes, the state needs %zu", bytes, state_bytes(h));
  const uint8_t* p = (const uint8_t*)buf + sizeof(h);
  // validate the path table before anything is changed
  for (size_t k = 0; k < e->paths.size(); k++) {
    StatePath sp;
    memcpy(&sp, p + k * sizeof(sp), sizeof(sp));
    BBX_REQUIRE(sp.cur >= -1 && sp.cur < (int)own.n_filters && sp.pend >= -1 && sp.pend < (int)own.n_filters,
                "bbx_engine_set_state: path %zu refers to a filter that does not exist", k);
  }
  DeviceGuard dg(e->device);
  int rc = drain(e);
  if (rc) return rc;
  for (size_t k = 0; k < e->paths.size(); k++, p += sizeof(StatePath)) {
    StatePath sp;
    memcpy(&sp, p, sizeof(sp));
    PathState& ps = e->paths[k];
    ps.cur = sp.cur >= 0 ? e->filters[sp.cur] : nullptr;
    ps.pend = sp.pend >= 0 ? e->filters[sp.pend] : nullptr;
    ps.has_pending = sp.has_pending != 0, ps.xfade = sp.xfade != 0;
    ps.gain = sp.gain, ps.delay = sp.delay, ps.pend_delay = sp.pend_delay;
  }
  for (int i = 0; i < 2; i++, p += h.xin_bytes) BBX_CUDA_TRY(cudaMemcpy(e->xin[i], p, h.xin_bytes, cudaMemcpyHostToDevice));
  BBX_CUDA_TRY(cudaMemcpy(e->fdl, p, h.fdl_bytes, cudaMemcpyHostToDevice));
  p += h.fdl_bytes;
  BBX_CUDA_TRY(cudaMemcpy(e->ybuf, p, h.ybuf_bytes, cudaMemcpyHostToDevice));
  e->head = h.head, e->wpos = h.wpos, e->parity = h.parity, e->tprev = h.tprev;
  e->steady_dirty = e->route_dirty = e->tc_dirty = true;
  return BBX_OK;
}

int bbx_engine_io_trace(bbx_engine* e, uint32_t calls) {
  BBX_REQUIRE(e != nullptr, "null engine");
  DeviceGuard dg(e->device);
  BBX_CUDA_TRY(cudaStreamSynchronize(e->s_in));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->s_out));
  while (e->io_ev.size() < (size_t)6 * calls) {
    cudaEvent_t ev;
    BBX_CUDA_TRY(cudaEventCreate(&ev));
    e->io_ev.push_back(ev);
  }
  e->io_cap = calls;
  e->io_calls = 0;
  return BBX_OK;
}

int bbx_engine_io_trace_read(bbx_engine* e, float* ms, uint32_t cap, uint32_t* n) {
  BBX_REQUIRE(e != nullptr && ms != nullptr && n != nullptr, "bbx_engine_io_trace_read: null argument");
  DeviceGuard dg(e->device);
  BBX_CUDA_TRY(cudaStreamSynchronize(e->s_in));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  BBX_CUDA_TRY(cudaStreamSynchronize(e->s_out));
  const uint32_t m = (uint32_t)std::min<size_t>(cap, e->io_calls);
  for (uint32_t i = 0; i < m; i++)
    for (int j = 0; j < 6; j++) {
      ms[6 * i + j] = 0.f;
      if (cudaEventElapsedTime(&ms[6 * i + j], e->io_ev[0], e->io_ev[6 * i + j]) != cudaSuccess) {
        cudaGetLastError();  // an event of a side the call did not use (direct I/O) was never recorded
        ms[6 * i + j] = -1.f;
      }
    }
  *n = m;
  e->io_cap = e->io_calls = 0;
  return BBX_OK;
}

int bbx_engine_exchange_time(bbx_engine* e, float* total_ms, uint64_t* exchanges, uint64_t* bytes_sent) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (size_t i = 0; i + 1 < e->xchg_events_used; i += 2) {
    float ms = 0.f;
    BBX_CUDA_TRY(cudaEventElapsedTime(&ms, e->xchg_events[i], e->xchg_events[i + 1]));
    e->xchg_ms_total += ms;
  }
  e->xchg_events_used = 0;
  if (total_ms) *total_ms = (float)e->xchg_ms_total;
  if (exchanges) *exchanges = e->xchg_count;
  if (bytes_sent) *bytes_sent = e->xchg_bytes;
  return BBX_OK;
}

int bbx_engine_mac_time(bbx_engine* e, float* total_ms, uint64_t* launches, uint64_t* channel_blocks,
                        uint64_t* algorithmic_bytes) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (size_t i = 0; i + 1 < e->mac_events_used; i += 2) {
    float ms = 0.f;
    BBX_CUDA_TRY(cudaEventElapsedTime(&ms, e->mac_events[i], e->mac_events[i + 1]));
    e->mac_ms_total += ms;
  }
  e->mac_events_used = 0;
  if (total_ms) *total_ms = (float)e->mac_ms_total;
  if (launches) *launches = e->mac_launches;
  if (channel_blocks) *channel_blocks = e->mac_units;
  if (algorithmic_bytes) *algorithmic_bytes = e->mac_bytes;
  return BBX_OK;
}

int bbx_engine_set_tuning(bbx_engine* e, uint32_t ctas_per_sm, uint32_t l2_keep_16ths, uint32_t time_tile) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  apply_tuning(e, ctas_per_sm, l2_keep_16ths, time_tile);
  return BBX_OK;
}

int bbx_engine_set_mixdown_kernel(bbx_engine* e, int per_output) {
  BBX_REQUIRE(e != nullptr, "bbx_engine_set_mixdown_kernel: null engine");
  e->pcm_out_mix = per_output == 0;
  return BBX_OK;
}

int bbx_engine_set_direct_io(bbx_engine* e, size_t max_bytes) {
  BBX_REQUIRE(e != nullptr, "bbx_engine_set_direct_io: null engine");
  e->direct_io_max_bytes = max_bytes;
  return BBX_OK;
}
uint64_t bbx_engine_direct_calls(const bbx_engine* e) { return e ? e->direct_calls : 0; }

int bbx_engine_set_fused(bbx_engine* e, int enable, uint32_t max_partitions) {
  BBX_REQUIRE(e != nullptr, "bbx_engine_set_fused: null engine");
  e->fused_on = enable != 0;
  if (max_partitions) e->fused_max_rows = max_partitions;
  return BBX_OK;
}
uint64_t bbx_engine_fused_calls(const bbx_engine* e) { return e ? e->fused_calls : 0; }

int bbx_engine_set_comm(bbx_engine* e, bbx_comm* c) {
  BBX_REQUIRE(e != nullptr, "null engine");
  BBX_REQUIRE(e->mode == BBX_MODE_MIMO, "bbx_engine_set_comm: only the MIMO engine has a collective");
  BBX_REQUIRE(!c || ((uint32_t)comm_world(c) == e->sh_world && (uint32_t)comm_rank(c) == e->sh_rank),
              "bbx_engine_set_comm: communicator (rank %d of %d) does not match the engine (rank %u of %u)", comm_rank(c),
              comm_world(c), e->sh_rank, e->sh_world);
  DeviceGuard dg(e->device);
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (c && !e->sh_send) {
    // world == 1 with a communicator: the sharded code path on one GPU (tests)
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_send, sizeof(float2) * (size_t)e->n_out * e->Tmax * e->B));
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_recv, sizeof(float2) * (size_t)e->sh_nloc * e->Tmax * e->B));
    std::vector<uint32_t> view(3 * (size_t)e->n_out);
    for (uint32_t o = 0; o < e->n_out; o++) {
      view[o] = o;
      view[e->n_out + o] = 1;
      view[2 * (size_t)e->n_out + o] = kNoJob;
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->sh_view, sizeof(uint32_t) * view.size()));
    BBX_CUDA_TRY(cudaMemcpy(e->sh_view, view.data(), sizeof(uint32_t) * view.size(), cudaMemcpyHostToDevice));
  }
  e->comm = c;
  return BBX_OK;
}

int bbx_engine_peer_export(bbx_engine* e, uint8_t* handle64) {
  BBX_REQUIRE(e && handle64, "bbx_engine_peer_export: null argument");
  BBX_REQUIRE(e->mode == BBX_MODE_MIMO && e->sh_world > 1 && e->sh_world <= 16,
              "bbx_engine_peer_export: only the input-sharded MIMO engine (2..16 ranks) exchanges spectra");
  DeviceGuard dg(e->device);
  if (!e->px_mem) {
    e->px_half = (size_t)e->sh_nloc * e->sh_world * e->Tmax * e->B;
    e->px_flag_off = (2 * e->px_half * sizeof(float2) + 255) & ~(size_t)255;
    const size_t bytes = e->px_flag_off + 2 * sizeof(uint32_t) * e->sh_world;
    BBX_CUDA_TRY(cudaMalloc((void**)&e->px_mem, bytes));
    BBX_CUDA_TRY(cudaMemset(e->px_mem, 0, bytes));
    BBX_CUDA_TRY(cudaMalloc((void**)&e->px_done, sizeof(uint32_t)));
    BBX_CUDA_TRY(cudaMemset(e->px_done, 0, sizeof(uint32_t)));
    std::vector<uint32_t> view(3 * (size_t)e->n_out);
    for (uint32_t o = 0; o < e->n_out; o++) {
      const bool mine = o >= e->sh_o0 && o < e->sh_o0 + e->sh_nloc;
      view[o] = mine ? (o - e->sh_o0) * e->sh_world : 0u;
      view[e->n_out + o] = e->sh_world;
      view[2 * (size_t)e->n_out + o] = kNoJob;
    }
    BBX_CUDA_TRY(cudaMalloc((void**)&e->px_view, sizeof(uint32_t) * view.size()));
    BBX_CUDA_TRY(cudaMemcpy(e->px_view, view.data(), sizeof(uint32_t) * view.size(), cudaMemcpyHostToDevice));
    BBX_CUDA_TRY(cudaHostAlloc((void**)&e->px_status_h, sizeof(int), cudaHostAllocMapped));
    *e->px_status_h = 0;
    BBX_CUDA_TRY(cudaHostGetDevicePointer((void**)&e->px_status, e->px_status_h, 0));
    BBX_CUDA_TRY(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  BBX_CUDA_TRY(cudaIpcGetMemHandle(&h, e->px_mem));
  static_assert(sizeof(h) == BBX_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  memcpy(handle64, &h, sizeof(h));
  return BBX_OK;
}

int bbx_engine_peer_attach(bbx_engine* e, const uint8_t* handles) {
  BBX_REQUIRE(e && handles, "bbx_engine_peer_attach: null argument");
  BBX_REQUIRE(e->px_mem != nullptr, "bbx_engine_peer_attach: call bbx_engine_peer_export first");
  DeviceGuard dg(e->device);
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (uint32_t r = 0; r < e->sh_world; r++) {
    uint8_t* base = e->px_mem;
    if (r != e->sh_rank) {
      if (!e->px_peer[r]) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * BBX_PEER_HANDLE_BYTES, sizeof(h));
        BBX_CUDA_TRY(cudaIpcOpenMemHandle(&e->px_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
      }
      base = (uint8_t*)e->px_peer[r];
    }
    e->px_table.data[r] = (float2*)base;
    e->px_table.flags[r] = (uint32_t*)(base + e->px_flag_off);
  }
  e->px_on = true;
  return BBX_OK;
}

int bbx_engine_tensor_status(bbx_engine* e, uint64_t* launches, int* status) {
  BBX_REQUIRE(e != nullptr, "null engine");
  DeviceGuard dg(e->device);
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (launches) *launches = e->tc_launches;
  if (status) {
    *status = 0;
    if (e->tc_status_h) *status = *(volatile int*)e->tc_status_h;
  }
  return BBX_OK;
}

__global__ void k_touch(float* x) { x[threadIdx.x] += 1.0f; }

int bbx_probe_launch_sync(int device, uint32_t ncalls, uint32_t warmup, double* us) {
  BBX_REQUIRE(us && ncalls, "bbx_probe_launch_sync: null argument");
  int rc = require_device();
  if (rc) return rc;
  DeviceGuard dg(device);
  float* x = nullptr;
  cudaStream_t s = nullptr;
  BBX_CUDA_TRY(cudaMalloc((void**)&x, 32 * sizeof(float)));
  BBX_CUDA_TRY(cudaMemset(x, 0, 32 * sizeof(float)));
  BBX_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  for (uint32_t i = 0; i < warmup + ncalls; i++) {
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    k_touch<<<1, 32, 0, s>>>(x);
    const cudaError_t se = cudaStreamSynchronize(s);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (se != cudaSuccess) {
      set_error("bbx_probe_launch_sync: %s", cudaGetErrorString(se));
      cudaStreamDestroy(s);
      cudaFree(x);
      return BBX_ERR_CUDA;
    }
    if (i >= warmup) us[i - warmup] = (double)(t1.tv_sec - t0.tv_sec) * 1e6 + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-3;
  }
  cudaStreamDestroy(s);
  cudaFree(x);
  return BBX_OK;
}

int bbx_probe_fp32_tflops(int device, float seconds, float* burst, float* sustained) {
  BBX_REQUIRE(burst || sustained, "bbx_probe_fp32_tflops: null outputs");
  int rc = require_device();
  if (rc) return rc;
  BBX_CUDA_TRY(cudaSetDevice(device));
  const int grid = kNumSMs * 4, iters = 4096;
  const double flops = 8.0 * 16 * (double)iters * grid * 256;  // 16 complex MACs = 64 FMA = 128 flop per thread and iteration
  float2* out = nullptr;
  BBX_CUDA_TRY(cudaMalloc((void**)&out, sizeof(float2) * (size_t)grid * 256));
  cudaEvent_t e0, e1;
  BBX_CUDA_TRY(cudaEventCreate(&e0));
  BBX_CUDA_TRY(cudaEventCreate(&e1));
  const float2 h = make_float2(1.0001f, 0.0001f), x = make_float2(0.9999f, 0.0002f);
  k_fp32_probe<<<grid, 256>>>(out, 64, h, x);  // warm-up
  BBX_CUDA_TRY(cudaDeviceSynchronize());
  float best = 0.f;
  for (int r = 0; r < 5; r++) {  // burst: best of five isolated launches
    BBX_CUDA_TRY(cudaEventRecord(e0));
    k_fp32_probe<<<grid, 256>>>(out, iters, h, x);
    BBX_CUDA_TRY(cudaEventRecord(e1));
    BBX_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    BBX_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (ms > 0.f) best = std::max(best, (float)(flops / (ms * 1e-3) / 1e12));
  }
  if (burst) *burst = best;
  if (sustained) {  // back-to-back launches for `seconds`: the rate under the power cap
    const int n = std::max(1, (int)(seconds * 1e3f / 1.3f));
    BBX_CUDA_TRY(cudaEventRecord(e0));
    for (int r = 0; r < n; r++) k_fp32_probe<<<grid, 256>>>(out, iters, h, x);
    BBX_CUDA_TRY(cudaEventRecord(e1));
    BBX_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    BBX_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *sustained = ms > 0.f ? (float)(flops * n / (ms * 1e-3) / 1e12) : 0.f;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return BBX_OK;
}

int bbx_engine_tensor_trace(bbx_engine* e, uint64_t* out, uint32_t max_ctas) {
  BBX_REQUIRE(e != nullptr, "null engine");
  DeviceGuard dg(e->device);
  BBX_CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (!out) {  // enable (max_ctas > 0) or disable
    cudaFree(e->tc_trace);
    e->tc_trace = nullptr;
    e->tc_trace_ctas = 0;
    if (max_ctas) {
      BBX_CUDA_TRY(cudaMalloc((void**)&e->tc_trace, sizeof(unsigned long long) * 16 * (size_t)max_ctas));
      BBX_CUDA_TRY(cudaMemset(e->tc_trace, 0, sizeof(unsigned long long) * 16 * (size_t)max_ctas));
      e->tc_trace_ctas = max_ctas;
    }
    return BBX_OK;
  }
  BBX_REQUIRE(e->tc_trace && max_ctas <= e->tc_trace_ctas, "bbx_engine_tensor_trace: tracing is not enabled for that many CTAs");
  BBX_CUDA_TRY(cudaMemcpy(out, e->tc_trace, sizeof(unsigned long long) * 16 * (size_t)max_ctas, cudaMemcpyDeviceToHost));
  return BBX_OK;
}

int bbx_engine_flush_l2(bbx_engine* e, size_t bytes) {
  BBX_REQUIRE(e != nullptr, "null engine");
  bytes = (bytes + 15) & ~(size_t)15;
  DeviceGuard dg(e->device);
  if (e->flush_bytes < bytes) {
    cudaStreamSynchronize(e->stream);
    cudaFree(e->flush_buf);
    e->flush_buf = nullptr;
    BBX_CUDA_TRY(cudaMalloc((void**)&e->flush_buf, bytes));
    e->flush_bytes = bytes;
  }
  k_flush<<<kNumSMs * 8, 256, 0, e->stream>>>(e->flush_buf, bytes / 16);
  BBX_CUDA_TRY(cudaGetLastError());
  return BBX_OK;
}

}  // extern "C"
