/* bbx.h -- C ABI of the B200-native partitioned-convolution engine (libbbx.so).
 *
 * This is the drop-in boundary for the bbcat-dsp convolution path: plain C, opaque
 * handles, plain pointers and sizes.  Each entry point names the reference interface
 * it replaces (paths relative to the bbcat-dsp tree).  Where the reference code is
 * absent from the tree (BlockConvolver / Convolver / FFT, README:38-51) the entry
 * point implements SURVEY.md 8.A.
 *
 * Conventions
 *   - every function returning int returns BBX_OK (0) or a negative bbx_status;
 *     bbx_last_error() gives the message of the last failure on the calling thread.
 *   - "host" entry points take host pointers and copy through the GPU; "_dev" entry
 *     points take device pointers on the current device and run on `stream`
 *     (a cudaStream_t passed as void*, NULL = the legacy default stream).
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with BBX_ERR_CUDA.
 *   - sample formats carry the numeric values of SampleFormat_t
 *     (src/SoundFormatConversions.h:20-37).
 */
#ifndef BBX_H
#define BBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBX_VERSION 0x000100

typedef enum {
  BBX_OK = 0,
  BBX_ERR_INVALID = -1,  /* bad argument / geometry */
  BBX_ERR_CUDA = -2,     /* CUDA runtime failure (message holds the CUDA error string) */
  BBX_ERR_NOMEM = -3,
  BBX_ERR_STATE = -4,    /* call not valid in the object's current state */
  BBX_ERR_UNSUPPORTED = -5
} bbx_status;

/* SampleFormat_t, src/SoundFormatConversions.h:20-37 (same numeric values) */
typedef enum {
  BBX_FMT_UNKNOWN = 0,
  BBX_FMT_16BIT = 1,
  BBX_FMT_24BIT = 2,
  BBX_FMT_32BIT = 3,
  BBX_FMT_FLOAT = 4,
  BBX_FMT_DOUBLE = 5,
  BBX_FMT_COUNT = 6
} bbx_sample_format;

/* ------------------------------------------------------------------------------------------
 * library
 * ---------------------------------------------------------------------------------------- */
int bbx_version(void);
const char* bbx_last_error(void);
/* number of visible CUDA devices (0 and BBX_ERR_CUDA when the runtime cannot initialise) */
int bbx_device_count(int* count);
/* pinned host memory for bbx_process() I/O buffers (cudaHostAlloc / cudaFreeHost) */
int bbx_host_alloc(void** ptr, size_t bytes);
int bbx_host_free(void* ptr);
/* block-contiguous channel shard of rank `rank` of `world`: pure host arithmetic */
int bbx_shard_range(uint32_t nchannels, uint32_t rank, uint32_t world, uint32_t* first, uint32_t* count);

/* ------------------------------------------------------------------------------------------
 * a1/a2  GetBitsPerSample / GetBytesPerSample / BlockTransferSanityChecks
 *        src/SoundFormatConversions.cpp:14-40, :59-93          (host integer logic, bit-exact)
 * ---------------------------------------------------------------------------------------- */
uint8_t bbx_get_bits_per_sample(int format);
uint8_t bbx_get_bytes_per_sample(int format);
/* returns 1 when the (clamped) transfer is non-empty and valid, else 0 */
int bbx_block_transfer_sanity_checks(uint32_t* src_channel, uint32_t* src_channels, uint32_t* dst_channel,
                                     uint32_t* dst_channels, uint32_t* nchannels, uint32_t* nframes,
                                     int allowsinglechannel);

/* ------------------------------------------------------------------------------------------
 * a3-a7  TransferSamples / TransferSamplesLinear
 *        src/SoundFormatConversions.cpp:151-198, :204-219 and the 2x2x6x6 converter table
 *        src/SoundFormatRawConversions.cpp:4516-4869 (all 100 non-NULL entries), ditherer == NULL (a7 below).
 *        Invalid geometry or an Unknown format is a silent no-op returning BBX_OK, like the
 *        reference (.cpp:160-166).  dst == src is allowed for the host form and gives the
 *        out-of-place result; any other overlap is undefined (SoundFormatConversions.h:128-130).
 * ---------------------------------------------------------------------------------------- */
int bbx_transfer_samples(const void* src, int srctype, int src_be, uint32_t src_channel, uint32_t src_channels,
                         void* dst, int dsttype, int dst_be, uint32_t dst_channel, uint32_t dst_channels,
                         uint32_t nchannels, uint32_t nframes);
int bbx_transfer_samples_dev(const void* src, int srctype, int src_be, uint32_t src_channel, uint32_t src_channels,
                             void* dst, int dsttype, int dst_be, uint32_t dst_channel, uint32_t dst_channels,
                             uint32_t nchannels, uint32_t nframes, void* stream);
int bbx_transfer_samples_linear(const void* src, int srctype, void* dst, int dsttype, uint32_t nsamples);

/* a7  Ditherer (src/SoundFormatConversions.h:39-54).  The reference calls ditherer->Dither(frame, sval, bits) between the
 * load and the quantisation of the narrowing converters only (src/SoundFormatRawConversions.cpp, e.g. :301, :749):
 * 24/32-bit -> 16-bit and float/double -> 16-bit with bits = 16, 32-bit/float/double -> 24-bit with bits = 8,
 * double -> 32-bit with bits = 0.  bbx_dither_bits returns that bit count, or -1 for a converter without a call site.
 * The tree ships only the no-op base class; arbitrary subclasses are honoured on the host by the C++ shim
 * (host/SoundFormatConversions.h: load -> hook -> quantise through the two entry points above).  Dither_TPDF -- named
 * by the reference's Dither_t enum, implemented nowhere in the tree -- is offered as a device option: triangular noise
 * of +-1 LSB of the destination word plus half an LSB before the truncating quantiser, from a counter-based generator
 * keyed by (seed, index of the sample in the transfer rectangle); DESIGN.md states the law, oracle/formats.c restates
 * it (parity unpinned against BBC output: there is none).  dither == BBX_DITHER_NONE is bbx_transfer_samples. */
typedef enum { BBX_DITHER_NONE = 0, BBX_DITHER_TPDF = 1 } bbx_dither; /* Dither_t, same numeric values */
int bbx_dither_bits(int srctype, int dsttype);
int bbx_transfer_samples_dither(const void* src, int srctype, int src_be, uint32_t src_channel, uint32_t src_channels,
                                void* dst, int dsttype, int dst_be, uint32_t dst_channel, uint32_t dst_channels,
                                uint32_t nchannels, uint32_t nframes, int dither, uint64_t seed);
int bbx_transfer_samples_dither_dev(const void* src, int srctype, int src_be, uint32_t src_channel, uint32_t src_channels,
                                    void* dst, int dsttype, int dst_be, uint32_t dst_channel, uint32_t dst_channels,
                                    uint32_t nchannels, uint32_t nframes, int dither, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------
 * a8/a9  MixSamples<T> (src/SoundMixing.h:55-81) and MixSamples(..., Interpolator&, inc)
 *        (src/SoundMixing.cpp:23-52, src/Interpolator.h:12-78).
 *        interp_state = {target, current}; it is advanced nframes steps like the caller's
 *        Interpolator object.  Products and sums are rounded separately (no FMA), so results
 *        equal the reference build bit for bit.
 * ---------------------------------------------------------------------------------------- */
int bbx_mix_samples_f32(const float* src, uint32_t src_channel, uint32_t src_channels, float* dst,
                        uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes, float mul);
int bbx_mix_samples_f64(const double* src, uint32_t src_channel, uint32_t src_channels, double* dst,
                        uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes, double mul);
int bbx_mix_samples_interp(const float* src, uint32_t src_channel, uint32_t src_channels, float* dst,
                           uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes,
                           float* interp_state, float inc);
int bbx_mix_samples_f32_dev(const float* src, uint32_t src_channel, uint32_t src_channels, float* dst,
                            uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes,
                            float mul, void* stream);
int bbx_mix_samples_interp_dev(const float* src, uint32_t src_channel, uint32_t src_channels, float* dst,
                               uint32_t dst_channel, uint32_t dst_channels, uint32_t nchannels, uint32_t nframes,
                               float* interp_state_host, float inc, void* stream);
/* Interpolator::operator+= applied nsteps times (src/Interpolator.h:55); host arithmetic */
int bbx_interpolator_step(float* interp_state, float inc, uint32_t nsteps);

/* ------------------------------------------------------------------------------------------
 * a10  FractionalSample (float / double buffers) and FractionalSampleAdditionalDelayRequired
 *      src/FractionalSample.cpp:249-341.  Batched: out[i] = FractionalSample(buffer, channel,
 *      channels, length, pos[i]).  Double accumulation in tap order, bit-exact.
 * ---------------------------------------------------------------------------------------- */
uint32_t bbx_fractional_sample_additional_delay_required(void);
int bbx_fractional_samples_f32(const float* buffer, uint32_t channel, uint32_t channels, uint32_t length,
                               const double* pos, uint32_t n, double* out);
int bbx_fractional_samples_f64(const double* buffer, uint32_t channel, uint32_t channels, uint32_t length,
                               const double* pos, uint32_t n, double* out);
int bbx_fractional_samples_f32_dev(const float* buffer, uint32_t channel, uint32_t channels, uint32_t length,
                                   const double* pos, uint32_t n, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a11  SoundDelayBuffer  src/SoundDelayBuffer.h:23-95, src/SoundDelayBuffer.cpp:11-191
 *      The ring lives in HBM, interleaved [length][channels] in `format`; src/dst are host
 *      pointers.  Same clamping, wrap-split and return values as the reference.
 * ---------------------------------------------------------------------------------------- */
typedef struct bbx_delay bbx_delay;
int bbx_delay_create(bbx_delay** out);
int bbx_delay_destroy(bbx_delay* d);
int bbx_delay_set_size(bbx_delay* d, uint32_t channels, uint32_t length, int format);
uint32_t bbx_delay_get_channels(const bbx_delay* d);
uint32_t bbx_delay_get_length(const bbx_delay* d);
uint32_t bbx_delay_get_write_position(const bbx_delay* d);
int bbx_delay_get_format(const bbx_delay* d);
uint32_t bbx_delay_write_samples(bbx_delay* d, const void* src, int srcformat, uint32_t channel, uint32_t nchannels,
                                 uint32_t nframes);
int bbx_delay_increment_write_position(bbx_delay* d, uint32_t nframes);
uint32_t bbx_delay_read_samples(bbx_delay* d, void* dst, int dstformat, uint32_t delay, uint32_t channel,
                                uint32_t nchannels, uint32_t nframes);
/* SoundDelayBuffer::ReadSample (.cpp:176-191) with the channel offset applied in SAMPLES */
float bbx_delay_read_sample(bbx_delay* d, uint32_t channel, uint32_t delay);
/* GetBuffer(): device pointer to the ring (format-native), NULL before SetSize */
const void* bbx_delay_get_buffer_dev(const bbx_delay* d);
/* copy the raw ring to host memory; returns bytes copied (0 on error) */
uint32_t bbx_delay_copy_buffer(const bbx_delay* d, void* dst, uint32_t maxbytes);
/* SoundRingBuffer  src/SoundDelayBuffer.h:105-181, src/SoundDelayBuffer.cpp:195-304: a SoundDelayBuffer with a read
 * position.  Create with bbx_ring_create and use the bbx_delay_* entry points above: like the reference's virtual
 * overrides they then limit SetSize / WriteSamples / IncrementWritePosition / ReadSamples by the read position (one
 * frame always stays free; reads stay relative to the WRITE position, as in the reference). */
int bbx_ring_create(bbx_delay** out);
uint32_t bbx_ring_get_read_position(const bbx_delay* d);
uint32_t bbx_ring_get_read_frames_available(const bbx_delay* d);
uint32_t bbx_ring_get_write_frames_available(const bbx_delay* d);
int bbx_ring_increment_read_position(bbx_delay* d, uint32_t nframes);

/* ------------------------------------------------------------------------------------------
 * a12-a14  BlockConvolver / Convolver / FFT  (absent from the tree: README:38-51, 68-69;
 *          behaviour = SURVEY.md 8.A)
 *
 * One engine = one multichannel convolver on one GPU.
 *   PER_CHANNEL  path c: input c -> filter -> delay -> output c            (n_paths = n_inputs)
 *   ROUTED       path p: input in(p) -> filter -> delay -> gain -> time-domain mix into out(p)
 *   MIMO         path o*n_inputs+i: frequency-domain sum over i into output o, one C2R per output
 * ---------------------------------------------------------------------------------------- */
typedef enum { BBX_MODE_PER_CHANNEL = 0, BBX_MODE_ROUTED = 1, BBX_MODE_MIMO = 2 } bbx_mode;

typedef struct {
  int device;               /* CUDA device ordinal */
  uint32_t block_size;      /* B: partition = block size, power of two, 64..4096 */
  uint32_t max_partitions;  /* longest filter in partitions (P_max) */
  uint32_t n_inputs;
  uint32_t n_outputs;       /* PER_CHANNEL: ignored (= n_inputs) */
  uint32_t n_paths;         /* ROUTED only; PER_CHANNEL = n_inputs, MIMO = n_inputs*n_outputs */
  int mode;                 /* bbx_mode */
  uint32_t max_blocks;      /* T_max: largest nframes/B accepted by one bbx_process call (>= 1) */
  uint32_t max_delay;       /* ceil of the largest per-path delay in samples */
  int fractional_delay;     /* 0: integer delays; 1: every delayed read goes through FractionalSample */
  uint32_t ring_length;     /* delay ring length R per path in frames; 0 = choose (see _get_ring_length) */
  /* tuning knobs; 0 = library default */
  uint32_t mac_ctas_per_sm;   /* streaming MAC: resident CTAs per SM, 1..4 (default 1: 148 row ranges) */
  uint32_t mac_l2_keep_16ths; /* streaming MAC: sixteenths of the H/FDL lines kept L2-resident, 1..16 (default 3); > 16 = hints off */
  uint32_t mac_time_tile;     /* 16 or 32: calls with >= tile/2 blocks and filters of >= 2*tile partitions use the time-batched MAC (default 16); 1 = streaming MAC only */
  uint32_t mimo_tensor;       /* MIMO mode: 0 = calls of >= 16 blocks run the per-bin complex GEMM on the tensor cores (3xTF32, tcgen05); 1 = SIMT MAC only */
  /* MIMO, input-sharded over several GPUs: this engine holds n_inputs of the inputs and ALL n_outputs
   * outputs of the matrix; after the MAC the partial output spectra are summed over the ranks with one NCCL
   * reduce-scatter per call and rank r converts outputs [r n_outputs / world, (r+1) n_outputs / world) to PCM
   * (bbx_process writes n_outputs / world channels).  0 or 1 = not sharded.  Needs bbx_engine_set_comm(). */
  uint32_t mimo_shard_world;
  uint32_t mimo_shard_rank;
  uint32_t reserved[2];
} bbx_config;

typedef struct bbx_engine bbx_engine;
typedef struct bbx_filter bbx_filter;

int bbx_engine_create(const bbx_config* cfg, bbx_engine** out);
int bbx_engine_destroy(bbx_engine* e);
/* R actually used for the per-path delay rings (the oracle must be run with the same R) */
uint32_t bbx_engine_get_ring_length(const bbx_engine* e);
/* stream the engine runs on (cudaStream_t as void*) */
void* bbx_engine_get_stream(const bbx_engine* e);

/* Filter object: H[p] = R2C_2B([h[pB .. pB+B-1], 0^B]), built on the device once, immutable,
 * shareable between paths of the engine that created it.
 * Ownership: the engine keeps a registry of its filters.  bbx_filter_destroy refuses (BBX_ERR_STATE) while a path still
 * has the filter selected or latched -- the MAC plans hold device pointers into its spectra; select another filter (or
 * NULL) and process one call first.  bbx_engine_destroy releases every filter the engine still owns; destroying such a
 * handle afterwards is a harmless no-op. */
int bbx_filter_create(bbx_engine* e, const float* ir, uint32_t length, bbx_filter** out);
int bbx_filter_destroy(bbx_filter* f);
uint32_t bbx_filter_partitions(const bbx_filter* f);
/* the filter object's spectra as the MAC kernels read them: [partitions][B] interleaved complex fp32, partition p =
 * R2C_2B([h[pB .. pB+B-1], 0^B]) / 2B with bin 0 holding (DC, Nyquist) (both real); for inspection and the FFT unit tests */
int bbx_filter_read_spectra(const bbx_filter* f, float* out, size_t max_floats);

/* ROUTED mode: connect path -> (input, output, gain).  Takes effect at the next bbx_process. */
int bbx_set_route(bbx_engine* e, uint32_t path, uint32_t input, uint32_t output, float gain);
/* Convolver::SelectFilter: latched, applied at the next block boundary (= first block of the
 * next bbx_process call); filter and delay switch together; crossfade != 0 blends old and new
 * over that one block with g_n = n/B.  filter == NULL silences the path. */
int bbx_set_filter(bbx_engine* e, uint32_t path, const bbx_filter* filter, int crossfade, double delay_samples);
/* the same for n paths in one call (a renderer that re-selects every channel's IR at once: C4); all n requests are validated
 * before any is latched; crossfade and delays may be NULL (= 0) */
int bbx_set_filters(bbx_engine* e, uint32_t n, const uint32_t* paths, const bbx_filter* const* filters, const int* crossfade,
                    const double* delays);

/* Process nframes (a multiple of B, at most max_blocks*B) of interleaved PCM.
 * in: [nframes][in_channels] in `infmt`, channels 0..n_inputs-1 are used.
 * out: [nframes][out_channels] in `outfmt`, channels 0..n_outputs-1 are written.
 * Host form: synchronous, includes H2D and D2H copies (pinned buffers avoid staging).
 * _dev form: device pointers, asynchronous on the engine stream. */
int bbx_process(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt,
                int out_be, uint32_t out_channels, uint32_t nframes);
int bbx_process_dev(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out,
                    int outfmt, int out_be, uint32_t out_channels, uint32_t nframes);
/* measurement aid: warmup + ncalls synchronous bbx_process calls on the same buffers, the host time of each of the last
 * ncalls (CLOCK_MONOTONIC around the call, microseconds) into us[ncalls] -- the per-block latency a C / C++ host sees at this
 * boundary, without the cost of a scripting language's foreign-function call around it */
int bbx_block_latency(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out, int outfmt, int out_be,
                      uint32_t out_channels, uint32_t nframes, uint32_t ncalls, uint32_t warmup, double* us);
/* Host pointers, asynchronous: returns once the H2D copy, the kernels and the D2H copy are enqueued (copies on
 * their own streams, three staging buffers per direction), so consecutive calls overlap transfer and compute.  `in` and
 * `out` must stay valid and untouched until bbx_engine_sync(); use pinned buffers (bbx_host_alloc). */
int bbx_process_async(bbx_engine* e, const void* in, int infmt, int in_be, uint32_t in_channels, void* out,
                      int outfmt, int out_be, uint32_t out_channels, uint32_t nframes);
int bbx_engine_sync(bbx_engine* e);

/* Communicator for the input-sharded MIMO engine: a thin handle over an NCCL communicator (libnccl is loaded at run
 * time).  Rank 0 creates the 128-byte id and distributes it out of band; every rank then calls bbx_comm_create. */
typedef struct bbx_comm bbx_comm;
int bbx_comm_available(void); /* 1 when libnccl.so.2 can be loaded */
int bbx_comm_unique_id(uint8_t* id128);
int bbx_comm_create(int world, int rank, const uint8_t* id128, int device, bbx_comm** out);
int bbx_comm_destroy(bbx_comm* c);
/* attach the communicator (world and rank must match bbx_config::mimo_shard_*); the engine does not own it */
int bbx_engine_set_comm(bbx_engine* e, bbx_comm* c);
/* Peer-memory mixdown: the same reduce without a collective library, for one process per GPU on an NVLink / NVSwitch box.
 * Every rank stores the partial spectrum of output o straight into the memory of the rank that owns o (peer stores into a
 * buffer shared through CUDA IPC) and the owner's inverse-transform kernel adds the `world` partials of an output in rank
 * order, so the sums do not depend on a collective's schedule; completion travels as epoch flags next to the data.
 * Set-up: every rank calls _peer_export, the application gathers the world handles in rank order (the same out-of-band
 * channel that carries the NCCL id) and every rank calls _peer_attach with all of them.  After that bbx_process* uses the
 * peer path instead of a communicator.  Crossfaded switches stay rejected; a rank that never arrives is reported by
 * bbx_engine_sync after a time-out. */
#define BBX_PEER_HANDLE_BYTES 64
int bbx_engine_peer_export(bbx_engine* e, uint8_t* handle64);
int bbx_engine_peer_attach(bbx_engine* e, const uint8_t* handles /* [world][BBX_PEER_HANDLE_BYTES] */);

/* Single-channel convenience = BlockConvolver::Convolve (README:38-39): path 0 of a
 * PER_CHANNEL engine, one float block in, one float block out (host pointers). */
int bbx_blockconvolver_convolve(bbx_engine* e, const float* in, float* out);

/* ------------------------------------------------------------------------------------------
 * next row  MultilayerBuffer<float>  src/MultilayerBuffer.h:19-431
 *      Output collection bus in HBM for renderers with different block sizes: every layer mixes
 *      (MixSamples) its blocks at its own write position; frames written by ALL layers can be read
 *      (TransferSamples when overwrite != 0, MixSamples otherwise) and are then shifted out.
 *      src / dst are host pointers.
 * ---------------------------------------------------------------------------------------- */
typedef struct bbx_mlb bbx_mlb;
int bbx_mlb_create(uint32_t channels, uint32_t layers, bbx_mlb** out);
int bbx_mlb_destroy(bbx_mlb* m);
uint32_t bbx_mlb_get_channels(const bbx_mlb* m);
uint32_t bbx_mlb_get_layers(const bbx_mlb* m);
uint32_t bbx_mlb_get_available_frames(const bbx_mlb* m);
int bbx_mlb_write_layer(bbx_mlb* m, uint32_t layer, const float* src, uint32_t srcchannel, uint32_t nsrcchannels,
                        uint32_t dstchannel, uint32_t nchannels, uint32_t nframes);
uint32_t bbx_mlb_read_buffer(bbx_mlb* m, uint32_t srcchannel, float* dst, uint32_t dstchannel, uint32_t ndstchannels,
                             uint32_t nchannels, uint32_t nframes, int overwrite);

/* ------------------------------------------------------------------------------------------
 * next row  BiQuadCoeffs / BiQuad  src/BiQuad.h:27-245, src/BiQuad.cpp:11-497
 *      One coefficient object shared by a bank of per-channel filters (what BiQuadFilterBank::Process runs per
 *      filter): coefficient design (CalcCoeffs) and explicit coefficients (SetCoeffs) with the reference's ramp
 *      (one Interpolate() step per frame), direct-form-II-transposed recurrence with double state, float samples,
 *      interleaved [frame][channel], dst == src allowed.  Bit-exact against the reference build.
 * ---------------------------------------------------------------------------------------- */
/* BiQuadCoeffs::Filter_t, src/BiQuad.h:31-42 (same numeric values) */
typedef enum {
  BBX_BIQUAD_FLAT = 0, BBX_BIQUAD_LPF6 = 1, BBX_BIQUAD_HPF6 = 2, BBX_BIQUAD_LPF12 = 3, BBX_BIQUAD_HPF12 = 4,
  BBX_BIQUAD_BPF = 5, BBX_BIQUAD_NOTCH = 6, BBX_BIQUAD_PEQ = 7, BBX_BIQUAD_LSH = 8, BBX_BIQUAD_HSH = 9
} bbx_biquad_type;
typedef struct bbx_biquad bbx_biquad;
/* normalised {num0, num1, num2, den1, den2} of a filter description (host arithmetic, no GPU needed) */
int bbx_biquad_calc_coeffs(int type, double freq, double fs, double gain, double bandwidth, double* out5);
int bbx_biquad_create(uint32_t nchannels, bbx_biquad** out);
int bbx_biquad_destroy(bbx_biquad* b);
/* BiQuadCoeffs::SetCoeffs: interp_samples > 0 ramps to the new coefficients over that many SAMPLES, else jumps */
int bbx_biquad_set_coeffs(bbx_biquad* b, const double* c5, double interp_samples);
/* BiQuadCoeffs::CalcCoeffs: interp_time in SECONDS */
int bbx_biquad_calc(bbx_biquad* b, int type, double freq, double fs, double gain, double bandwidth, double interp_time);
/* BiQuad::Process(filters, src, dst, nchannels, nsrcchannels, ndstchannels, nframes, coeffs); host pointers */
int bbx_biquad_process(bbx_biquad* b, const float* src, float* dst, uint32_t nchannels, uint32_t nsrcchannels,
                       uint32_t ndstchannels, uint32_t nframes);
int bbx_biquad_process_dev(bbx_biquad* b, const float* src, float* dst, uint32_t nchannels, uint32_t nsrcchannels,
                           uint32_t ndstchannels, uint32_t nframes, void* stream);
/* filter states w[nchannels][2], current coefficients, {mul, dec} of the ramp (any pointer may be NULL) */
int bbx_biquad_get_state(const bbx_biquad* b, double* w, double* cur5, double* mul_dec);
int bbx_biquad_reset(bbx_biquad* b); /* BiQuad::Reset on every filter */

/* BiQuadFilterBank  src/BiQuad.h:247-353, src/BiQuad.cpp:498-662: nfilters biquads in series on each of nchannels
 * channels, one coefficient object (with its own ramp) per filter.  Process (src/BiQuad.cpp:639-662) is the reference's
 * filter-by-filter loop -- BiQuad::Process from src to dst for the first filter, in place on dst for the others -- fused
 * into ONE pass: a thread takes a frame of its channel through all filters before the next frame (one read of src, one
 * write of dst; banks of more than 16 filters run in passes of 16).  Same IEEE operations in the same order per
 * (filter, channel): bit-exact against the reference build.  SetFilters / SetChannels semantics: filters are dropped
 * from / appended at the end, surviving (filter, channel) pairs keep their audio state, new ones start at zero with
 * flat coefficients; AddFilter appends a filter with the given coefficients (no ramp). */
typedef struct bbx_fbank bbx_fbank;
int bbx_fbank_create(uint32_t nchannels, uint32_t nfilters, bbx_fbank** out);
int bbx_fbank_destroy(bbx_fbank* f);
int bbx_fbank_set_filters(bbx_fbank* f, uint32_t n);           /* BiQuadFilterBank::SetFilters */
int bbx_fbank_add_filter(bbx_fbank* f, const double* c5);      /* BiQuadFilterBank::AddFilter */
int bbx_fbank_set_channels(bbx_fbank* f, uint32_t n);          /* BiQuadFilterBank::SetChannels */
int bbx_fbank_get_size(const bbx_fbank* f, uint32_t* nchannels, uint32_t* nfilters); /* GetChannels / GetFilters */
/* GetFilterCoeffs(filter)->SetCoeffs / ->CalcCoeffs (interp_samples in SAMPLES, interp_time in SECONDS) */
int bbx_fbank_set_coeffs(bbx_fbank* f, uint32_t filter, const double* c5, double interp_samples);
int bbx_fbank_calc(bbx_fbank* f, uint32_t filter, int type, double freq, double fs, double gain, double bandwidth,
                   double interp_time);
/* BiQuadFilterBank::Process(src, dst, nchannels, nsrcchannels, ndstchannels, nframes); host pointers / device pointers */
int bbx_fbank_process(bbx_fbank* f, const float* src, float* dst, uint32_t nchannels, uint32_t nsrcchannels,
                      uint32_t ndstchannels, uint32_t nframes);
int bbx_fbank_process_dev(bbx_fbank* f, const float* src, float* dst, uint32_t nchannels, uint32_t nsrcchannels,
                          uint32_t ndstchannels, uint32_t nframes, void* stream);
/* state of one filter: w[nchannels][2], current coefficients, {mul, dec} of its ramp (any pointer may be NULL) */
int bbx_fbank_get_state(const bbx_fbank* f, uint32_t filter, double* w, double* cur5, double* mul_dec);
int bbx_fbank_reset(bbx_fbank* f);                             /* BiQuadFilterBank::Reset */
int bbx_fbank_launches(const bbx_fbank* f, uint64_t* launches); /* kernels launched so far (one per 16 filters per call) */

/* BiQuadCascade  src/BiQuad.h:373-792: a bank of nchannels independent cascades of numfilters (1..12) biquads in float,
 * both forms of Tick -- the plain cascade and the "vectorised" pipeline of the SSE3 build (numfilters a multiple of four,
 * else switched off like the reference; it delays the signal by numfilters - 1 samples).  unroll is accepted for signature
 * parity (same arithmetic).  The output gain g of the coefficient vector is stored and never applied, as in the reference.
 * More than 12 filters is an error here (the reference logs and leaves an unusable object).  Bit-exact. */
typedef struct bbx_cascade bbx_cascade;
int bbx_cascade_create(uint32_t nchannels, uint32_t numfilters, int vectorise, int unroll, bbx_cascade** out);
int bbx_cascade_destroy(bbx_cascade* c);
/* BiQuadCascade::SetCoefficients(const std::vector<float>&): (g, b1[0], b2[0], a1[0], a2[0], b1[1], ...), n = 4 * numfilters + 1,
 * resets the registers; channel 0xFFFFFFFF = every cascade of the bank */
int bbx_cascade_set_coefficients(bbx_cascade* c, uint32_t channel, const float* coeffs, uint32_t n);
int bbx_cascade_reset(bbx_cascade* c); /* BiQuadCascade::Reset on every channel */
/* ProcessCascade on every channel; host buffers [nframes][nchannels] (interleaved != 0) or [nchannels][nframes] */
int bbx_cascade_process(bbx_cascade* c, const float* src, float* dst, uint32_t nframes, int interleaved);
/* device pointers; channel j reads src[j * src_channel_stride + i * src_frame_stride] (strides in samples) */
int bbx_cascade_process_dev(bbx_cascade* c, const float* src, long long src_channel_stride, long long src_frame_stride,
                            float* dst, long long dst_channel_stride, long long dst_frame_stride, uint32_t nframes,
                            void* stream);
/* registers x, y, w0, w1 (12 floats each) and lastoutput of one channel (any pointer may be NULL);
 * returns numfilters | vectorise << 8 */
uint32_t bbx_cascade_get_state(const bbx_cascade* c, uint32_t channel, float* x12, float* y12, float* w0_12, float* w1_12,
                               float* lastoutput);

/* AllPassFilterChain<float>  src/AllPassFilter.h:12-262 (ring semantics src/RingBuffer.h:17-121): nfilters Schroeder
 * all-pass sections (delay[f] >= 1 frames, coefficient[f]) over nchannels interleaved channels, rings in HBM in the
 * reference's layout; Process = AllPassFilterChain::Process (dst may equal src).  Bit-exact. */
typedef struct bbx_allpass bbx_allpass;
int bbx_allpass_create(uint32_t nchannels, uint32_t nfilters, const uint32_t* delays, const float* coeffs, bbx_allpass** out);
int bbx_allpass_destroy(bbx_allpass* a);
int bbx_allpass_process(bbx_allpass* a, const float* src, float* dst, uint32_t srcchannel, uint32_t nsrcchannels,
                        uint32_t dstchannel, uint32_t ndstchannels, uint32_t nframes);
int bbx_allpass_process_dev(bbx_allpass* a, const float* src, float* dst, uint32_t srcchannel, uint32_t nsrcchannels,
                            uint32_t dstchannel, uint32_t ndstchannels, uint32_t nframes, void* stream);
/* raw ring of section `filter` (nchannels * delay floats) into `ring`; returns RingBuffer::GetPosition() */
uint32_t bbx_allpass_get_state(const bbx_allpass* a, uint32_t filter, float* ring, uint32_t maxitems);

/* ------------------------------------------------------------------------------------------
 * measurement hooks (bench.py): CUDA-event timing on the engine stream, launch counting and
 * the dominant kernel's (FDL MAC) accumulated device time.
 * ---------------------------------------------------------------------------------------- */
int bbx_engine_timer_start(bbx_engine* e);
int bbx_engine_timer_stop(bbx_engine* e, float* elapsed_ms); /* synchronises the stream */
uint64_t bbx_engine_launch_count(const bbx_engine* e);
/* name of the multiply-accumulate kernel the most recent call ran (static string) */
const char* bbx_engine_mac_kernel(const bbx_engine* e);
/* when enabled every MAC launch is bracketed by events; _mac_time sums them (synchronises) */
int bbx_engine_profile_mac(bbx_engine* e, int enable);
int bbx_engine_mac_time(bbx_engine* e, float* total_ms, uint64_t* launches, uint64_t* channel_blocks,
                        uint64_t* algorithmic_bytes);
/* input-sharded MIMO engine, while _profile_mac is enabled: accumulated device time of the exchange step of every call
 * (k_gather_spectra_peer + k_peer_wait, or k_gather_spectra + ncclReduceScatter), the number of exchanges and the bytes
 * this rank sent to its peers (synchronises) */
int bbx_engine_exchange_time(bbx_engine* e, float* total_ms, uint64_t* exchanges, uint64_t* bytes_sent);
/* Checkpoint / resume (SURVEY.md aux): the audio state of an engine -- FDL ring, previous input block, delay rings, ring
 * positions, and per path the selected / latched filter, delays and gain -- as one host blob.  Filters are referred to by
 * their position among the engine's live filters, so the restoring engine (the same one later, or a fresh one of the same
 * bbx_config) must hold the same filters created in the same order.  Calls made after set_state produce the bytes the
 * saving engine produced after get_state.  Both calls drain the engine first.  Not available on input-sharded engines. */
int bbx_engine_state_size(const bbx_engine* e, size_t* bytes);
int bbx_engine_get_state(bbx_engine* e, void* buf, size_t capacity);
int bbx_engine_set_state(bbx_engine* e, const void* buf, size_t bytes);
/* Trace of the host-buffer pipeline: the next `calls` bbx_process_async calls record a timing event at the start and end
 * of their H2D copy, their kernels and their D2H copy; _read returns, per traced call, those six times in ms relative to
 * the first call's H2D start ([n][6]: h2d0, h2d1, kernels0, kernels1, d2h0, d2h1; -1 = that side was not used). */
int bbx_engine_io_trace(bbx_engine* e, uint32_t calls);
int bbx_engine_io_trace_read(bbx_engine* e, float* ms, uint32_t cap, uint32_t* n);
/* change the tuning knobs of bbx_config at run time (0 = leave as is); takes effect at the next call */
int bbx_engine_set_tuning(bbx_engine* e, uint32_t ctas_per_sm, uint32_t l2_keep_16ths, uint32_t time_tile);
/* mixdowns of many paths into few outputs (>= 4 paths per output, <= 32 outputs, <= 256 routes) run k_pcm_out_mix, which
 * forms the delayed, gained path samples of a tile in parallel and adds them in route order; per_output != 0 keeps them
 * on the per-output kernel (tuning / A-B: the bytes are identical); the same switch keeps fractional-delay engines on
 * the tile kernel instead of the one-sample-per-thread kernel k_pcm_out_frac */
int bbx_engine_set_mixdown_kernel(bbx_engine* e, int per_output);
/* latency path of bbx_process / bbx_process_async, decided for the input and the output side separately: a PCM buffer of
 * at most max_bytes in pinned host memory the device can address (bbx_host_alloc, cudaHostAlloc, cudaHostRegister) is
 * read / written by the PCM kernels directly over PCIe instead of going through a copy engine and the staging buffer.
 * Only layouts that suit the bus qualify: typed little-endian samples (not 24-bit) and at least 128 bytes of used
 * channels per frame.  Default 1 MiB; 0 disables (every host call is staged).  _direct_calls counts the calls in which
 * at least one side took this path. */
int bbx_engine_set_direct_io(bbx_engine* e, size_t max_bytes);
uint64_t bbx_engine_direct_calls(const bbx_engine* e);
/* single-launch latency path: a streaming call (one block) of a PER_CHANNEL or ROUTED engine whose paths have at most
 * max_partitions partitions (default 32) runs k_block_fused -- PCM in, forward transform, MAC, inverse transform,
 * crossfade, delay ring and, for PER_CHANNEL engines, the output stage in ONE launch (ROUTED engines add their mixdown
 * launch) -- instead of five dependent launches.  The bytes are identical to the multi-kernel path; enable = 0 keeps
 * every call on the multi-kernel path (A/B, tests), max_partitions = 0 leaves the limit as is.  _fused_calls counts. */
int bbx_engine_set_fused(bbx_engine* e, int enable, uint32_t max_partitions);
uint64_t bbx_engine_fused_calls(const bbx_engine* e);
/* tensor-core MIMO path: number of k_mimo_tc launches so far and the device status word (0 = ok; non-zero =
 * a barrier wait timed out inside the kernel, results invalid).  Synchronises the stream. */
int bbx_engine_tensor_status(bbx_engine* e, uint64_t* launches, int* status);
/* profiling hook of k_mimo_tc: out == NULL enables (max_ctas > 0) / disables the per-CTA role trace; otherwise copies
 * 16 cycle counters per CTA of the last launch: [0,1] loader total / waiting for a raw slot, [2..4] MMA issuer total /
 * waiting for operands / waiting for the read-out, [5,6] epilogue total / waiting for accumulators, [7+3g..9+3g]
 * producer group g total / waiting for raw data / waiting for its operand stage */
int bbx_engine_tensor_trace(bbx_engine* e, uint64_t* out, uint32_t max_ctas);
/* FP32 roofline probe: the rate (TFLOP/s) a pure packed-FMA kernel with the MAC's operand pattern reaches on this GPU:
 * best of five isolated launches (burst) and averaged over `seconds` of back-to-back launches (sustained, power cap) */
int bbx_probe_fp32_tflops(int device, float seconds, float* burst, float* sustained);
/* the floor under every per-block latency on this host: one trivial kernel on a stream + cudaStreamSynchronize, the host
 * time of each of ncalls round trips (after warmup) in microseconds, timed inside the library like bbx_block_latency */
int bbx_probe_launch_sync(int device, uint32_t ncalls, uint32_t warmup, double* us);
/* write `bytes` of a scratch buffer on the engine stream (L2 flush between timed iterations) */
int bbx_engine_flush_l2(bbx_engine* e, size_t bytes);

/* ---- SOFA (AES69) impulse-response sets: the on-disk side of IR selection ------------------------------------------
 * Replaces: src/SOFA.{h,cpp} ("SOFA file support via the netcdf-bbc libraries", README:77-78; libnetcdf dependency in
 * debian/control:5) -- listed by the reference, ABSENT from the mounted tree, so the interface below is ours and parity is
 * unpinned against BBC; the container parser is pinned against an independent netCDF implementation (scipy.io.netcdf_file,
 * tests/test_sofa.py).  Reads the netCDF classic container (CDF-1 / CDF-2) with the SOFA names: Data.IR [M][R][N] (DataType
 * FIR) or [M][R][E][N] (FIRE), Data.Delay [I|M][R]([E]) in samples, Data.SamplingRate [I|M], Source/ListenerPosition
 * [I|M][C], Receiver/EmitterPosition [R|E][C][I|M].  A netCDF-4 (HDF5) container is refused with BBX_ERR_UNSUPPORTED and a
 * message (no HDF5 implementation in this build).  Host-side objects: no device needed except for _create_filters. */
typedef struct bbx_sofa bbx_sofa;
enum { BBX_SOFA_SOURCE = 0, BBX_SOFA_LISTENER = 1, BBX_SOFA_RECEIVER = 2, BBX_SOFA_EMITTER = 3 };
int bbx_sofa_open(const char* path, bbx_sofa** out);
int bbx_sofa_open_memory(const void* data, size_t bytes, bbx_sofa** out);
int bbx_sofa_close(bbx_sofa* s);
/* M measurements, R receivers (ears), E emitters (1 for FIR), N samples per impulse response; any pointer may be NULL */
int bbx_sofa_get_sizes(const bbx_sofa* s, uint32_t* M, uint32_t* R, uint32_t* E, uint32_t* N);
int bbx_sofa_get_samplerate(const bbx_sofa* s, uint32_t measurement, double* hz);
/* impulse response (measurement, receiver, emitter) as fp32 (Data.IR is stored in double), n values: truncated or
 * zero-padded to n */
int bbx_sofa_get_ir(const bbx_sofa* s, uint32_t measurement, uint32_t receiver, uint32_t emitter, float* dst, uint32_t n);
/* Data.Delay of (measurement, receiver, emitter) in samples: the `delay_samples` of bbx_set_filter */
int bbx_sofa_get_delay(const bbx_sofa* s, uint32_t measurement, uint32_t receiver, uint32_t emitter, double* samples);
/* position variable `which` (BBX_SOFA_*): row `index` (a measurement for source / listener, an object otherwise), the three
 * coordinates as stored, *spherical = 1 when the variable's Type attribute says so (degrees, degrees, metres) */
int bbx_sofa_get_position(const bbx_sofa* s, int which, uint32_t index, double xyz[3], int* spherical);
/* the measurement whose SourcePosition is nearest to pos (Euclidean distance after conversion to cartesian; a spherical
 * query with radius <= 0 selects by direction only); ties go to the lowest index */
int bbx_sofa_nearest_measurement(const bbx_sofa* s, const double pos[3], int spherical, uint32_t* measurement);
/* global attribute as text (numeric attributes printed with %.17g) */
int bbx_sofa_get_attribute(const bbx_sofa* s, const char* name, char* buf, uint32_t buflen);
/* one filter object per measurement for (receiver, emitter): the IR bank a renderer selects from with bbx_set_filter;
 * count must equal M; on failure nothing is left allocated */
int bbx_sofa_create_filters(const bbx_sofa* s, bbx_engine* e, uint32_t receiver, uint32_t emitter, bbx_filter** out, uint32_t count);

#ifdef __cplusplus
}
#endif
#endif /* BBX_H */
