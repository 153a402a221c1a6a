/* TEST INFRASTRUCTURE ONLY -- the CPU oracle.  Nothing in the product path
 * (bbcat-dsp_b200/, include/) may include, link or call anything declared here;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do.
 *
 * Plain-C restatement of the bbcat-dsp partitioned-convolution path:
 *   - the in-tree pieces (formats / mixing / interpolator / fractional sample /
 *     delay ring) follow /root/reference/src line by line and are PINNED against
 *     the reference's own code compiled into oracle/_ref/libbbcref.so and against
 *     tests/golden/ *.npz generated from it (tools/gen_golden.py);
 *   - BlockConvolver / Convolver / FFT are ABSENT from the mounted reference
 *     (README:38-51 vs src/CMakeLists.txt:2-25) and FFTW (debian/control:5) is not
 *     installed, so upols.c / convolver.c implement SURVEY.md 8.A (normative) with an
 *     own FFT, and are pinned against a float64 direct convolution (direct.c) and
 *     numpy/scipy float64.  PARITY UNPINNED against any BBC convolver output: none
 *     exists in the tree (no tests, no golden vectors).
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SampleFormat_t numeric values, SoundFormatConversions.h:20-37 */
enum {
  ORC_FMT_UNKNOWN = 0,
  ORC_FMT_16BIT = 1,
  ORC_FMT_24BIT = 2,
  ORC_FMT_32BIT = 3,
  ORC_FMT_FLOAT = 4,
  ORC_FMT_DOUBLE = 5,
  ORC_FMT_COUNT = 6
};

/* ---- formats.c : SoundFormatConversions.cpp / SoundFormatRawConversions.cpp ---- */
unsigned orc_get_bits_per_sample(int fmt);
unsigned orc_get_bytes_per_sample(int fmt);
int orc_block_transfer_sanity_checks(unsigned* src_channel, unsigned* src_channels, unsigned* dst_channel,
                                     unsigned* dst_channels, unsigned* nchannels, unsigned* nframes,
                                     int allowsinglechannel);
void orc_transfer_samples(const void* src, int srctype, int src_be, unsigned src_channel, unsigned src_channels,
                          void* dst, int dsttype, int dst_be, unsigned dst_channel, unsigned dst_channels,
                          unsigned nchannels, unsigned nframes);
void orc_transfer_samples_linear(const void* src, int srctype, void* dst, int dsttype, unsigned nsamples);
/* a7 Ditherer (SoundFormatConversions.h:39-54): call sites and bit counts of the converter table; the hook form restates
 * where the reference calls ditherer->Dither, the TPDF form restates libbbx's own law (see formats.c) */
typedef void (*orc_dither_hook)(void* user, unsigned frame_counter, int kind, void* data, unsigned bits);
int orc_dither_bits(int srctype, int dsttype);
void orc_transfer_samples_dither(const void* src, int srctype, int src_be, unsigned src_channel, unsigned src_channels,
                                 void* dst, int dsttype, int dst_be, unsigned dst_channel, unsigned dst_channels,
                                 unsigned nchannels, unsigned nframes, int dither, uint64_t seed, orc_dither_hook hook,
                                 void* user);
/* the hook of tests/cpp/test_ditherer.h restated in C (user = uint32_t call counter) */
void orc_test_dither_hook(void* user, unsigned frame_counter, int kind, void* data, unsigned bits);
unsigned orc_transfer_samples_ditherer(const void* src, int srctype, int src_be, unsigned src_channel, unsigned src_channels,
                                       void* dst, int dsttype, int dst_be, unsigned dst_channel, unsigned dst_channels,
                                       unsigned nchannels, unsigned nframes, int mode);

/* ---- mixing.c : SoundMixing.h/.cpp, Interpolator.h ---- */
void orc_mix_samples_f32(const float* src, unsigned src_channel, unsigned src_channels, float* dst,
                         unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes, float mul);
void orc_mix_samples_f64(const double* src, unsigned src_channel, unsigned src_channels, double* dst,
                         unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes, double mul);
/* interp_state = {target, current}, advanced in place */
void orc_mix_samples_interp(const float* src, unsigned src_channel, unsigned src_channels, float* dst,
                            unsigned dst_channel, unsigned dst_channels, unsigned nchannels, unsigned nframes,
                            float* interp_state, float inc);
void orc_interpolator_step(float* interp_state, float inc, unsigned nsteps);

/* ---- fracsample.c : FractionalSample.cpp ---- */
unsigned orc_fractional_sample_additional_delay_required(void);
double orc_fractional_sample_f32(const float* buffer, unsigned channel, unsigned channels, unsigned length, double pos);
double orc_fractional_sample_f64(const double* buffer, unsigned channel, unsigned channels, unsigned length, double pos);
void orc_fractional_samples_f32(const float* buffer, unsigned channel, unsigned channels, unsigned length,
                                const double* pos, unsigned n, double* out);
void orc_fractional_samples_f64(const double* buffer, unsigned channel, unsigned channels, unsigned length,
                                const double* pos, unsigned n, double* out);

/* ---- delaybuf.c : SoundDelayBuffer.cpp:26-191 ---- */
typedef struct orc_delay orc_delay;
orc_delay* orc_delay_create(void);
void orc_delay_destroy(orc_delay* d);
void orc_delay_set_size(orc_delay* d, unsigned chans, unsigned length, int fmt);
unsigned orc_delay_get_channels(const orc_delay* d);
unsigned orc_delay_get_length(const orc_delay* d);
unsigned orc_delay_get_write_position(const orc_delay* d);
int orc_delay_get_format(const orc_delay* d);
unsigned orc_delay_write_samples(orc_delay* d, const void* src, int srcformat, unsigned channel, unsigned nchannels,
                                 unsigned nframes);
void orc_delay_increment_write_position(orc_delay* d, unsigned nframes);
unsigned orc_delay_read_samples(orc_delay* d, void* dst, int dstformat, unsigned delay, unsigned channel,
                                unsigned nchannels, unsigned nframes);
unsigned orc_delay_copy_buffer(const orc_delay* d, void* dst, unsigned maxbytes);
/* SoundRingBuffer (SoundDelayBuffer.h:105-181, .cpp:195-304): created with orc_ring_create, used through the
 * orc_delay_* entry points (which apply the ring's limits, like the reference's virtual methods) plus: */
orc_delay* orc_ring_create(void);
unsigned orc_ring_get_read_position(const orc_delay* d);
unsigned orc_ring_get_read_frames_available(const orc_delay* d);
unsigned orc_ring_get_write_frames_available(const orc_delay* d);
void orc_ring_increment_read_position(orc_delay* d, unsigned nframes);

/* ---- multilayer.c : MultilayerBuffer.h ("next" row) ---- */
typedef struct orc_mlb orc_mlb;
orc_mlb* orc_mlb_create(unsigned channels, unsigned layers);
void orc_mlb_destroy(orc_mlb* m);
void orc_mlb_write_layer(orc_mlb* m, unsigned layer, const float* src, unsigned srcchannel, unsigned nsrcchannels,
                         unsigned dstchannel, unsigned nchannels, unsigned nframes);
unsigned orc_mlb_available_frames(const orc_mlb* m);
unsigned orc_mlb_read_buffer(orc_mlb* m, unsigned srcchannel, float* dst, unsigned dstchannel, unsigned ndstchannels,
                             unsigned nchannels, unsigned nframes, int overwrite);

/* ---- biquad.c : BiQuadCoeffs / BiQuad ("next" row) ---- */
/* filter types: src/BiQuad.h:31-42 (FLAT 0, LPF6 1, HPF6 2, LPF12 3, HPF12 4, BPF 5, NOTCH 6, PEQ 7, LSH 8, HSH 9) */
typedef struct orc_biquad orc_biquad;
void orc_biquad_calc_coeffs(int type, double freq, double fs, double gain, double bandwidth, double* out5);
orc_biquad* orc_biquad_create(unsigned nch);
void orc_biquad_destroy(orc_biquad* b);
void orc_biquad_set_coeffs(orc_biquad* b, const double* c5, double interp_samples);
void orc_biquad_calc(orc_biquad* b, int type, double freq, double fs, double gain, double bandwidth, double interp_time);
void orc_biquad_process(orc_biquad* b, const float* src, float* dst, unsigned nchannels, unsigned nsrc, unsigned ndst,
                        unsigned nframes);
void orc_biquad_get_state(const orc_biquad* b, double* w, double* cur5, double* mul_dec);
void orc_biquad_reset(orc_biquad* b);
/* BiQuadFilterBank (src/BiQuad.cpp:498-662): the reference's filter-by-filter loop */
typedef struct orc_fbank orc_fbank;
orc_fbank* orc_fbank_create(unsigned nch, unsigned nfilters);
void orc_fbank_destroy(orc_fbank* b);
void orc_fbank_set_filters(orc_fbank* b, unsigned n);
void orc_fbank_add_filter(orc_fbank* b, const double* c5);
void orc_fbank_set_channels(orc_fbank* b, unsigned n);
void orc_fbank_set_coeffs(orc_fbank* b, unsigned filter, const double* c5, double interp_samples);
void orc_fbank_calc(orc_fbank* b, unsigned filter, int type, double freq, double fs, double gain, double bandwidth,
                    double interp_time);
void orc_fbank_process(orc_fbank* b, const float* src, float* dst, unsigned nchannels, unsigned nsrc, unsigned ndst,
                       unsigned nframes);
void orc_fbank_get_state(const orc_fbank* b, unsigned filter, double* w, double* cur5, double* mul_dec);
void orc_fbank_reset(orc_fbank* b);

/* ---- allpass.c : AllPassFilterChain<float> ("next" row) ---- */
typedef struct orc_allpass orc_allpass;
orc_allpass* orc_allpass_create(unsigned nchannels, unsigned nfilters, const unsigned* delays, const float* coeffs);
void orc_allpass_destroy(orc_allpass* a);
void orc_allpass_process(orc_allpass* a, const float* src, float* dst, unsigned srcchannel, unsigned nsrc, unsigned dstchannel,
                         unsigned ndst, unsigned nframes);
unsigned orc_allpass_get_state(const orc_allpass* a, unsigned f, float* ring, unsigned maxitems);

/* ---- cascade.c : BiQuadCascade, one per channel ("next" row) ---- */
typedef struct orc_cascade orc_cascade;
orc_cascade* orc_cascade_create(unsigned nchannels, unsigned numfilters, int vectorise, int unroll);
void orc_cascade_destroy(orc_cascade* b);
/* channel == ~0u: every channel; n must be 4 * numfilters + 1; returns 1 on success */
int orc_cascade_set_coefficients(orc_cascade* b, unsigned channel, const float* coeffs, unsigned n);
void orc_cascade_reset(orc_cascade* b);
/* channel j reads src[j * src_cs + i * src_fs], writes dst[j * dst_cs + i * dst_fs] */
void orc_cascade_process(orc_cascade* b, const float* src, long src_cs, long src_fs, float* dst, long dst_cs, long dst_fs,
                         unsigned nframes);
/* registers of one channel (12 floats each); returns numfilters | vectorise << 8 */
unsigned orc_cascade_get_state(const orc_cascade* b, unsigned channel, float* x12, float* y12, float* w0_12, float* w1_12,
                               float* last);

/* ---- fft.c : own FFT (FFTW stand-in, unnormalised both directions) ---- */
/* complex in-place FFT of n (power of two) interleaved float pairs; inverse != 0 conjugates the kernel */
void orc_cfft(float* data, unsigned n, int inverse);
/* real -> half-complex: in[n] real, out[n/2+1] interleaved complex */
void orc_rfft(const float* in, float* out, unsigned n);
/* half-complex -> real, unnormalised (caller scales by 1/n) */
void orc_irfft(const float* in, float* out, unsigned n);

/* ---- upols.c : SURVEY.md 8.A BlockConvolver ---- */
typedef struct orc_filter orc_filter;
orc_filter* orc_filter_create(const float* ir, unsigned length, unsigned block);
void orc_filter_destroy(orc_filter* f);
unsigned orc_filter_partitions(const orc_filter* f);
unsigned orc_filter_block(const orc_filter* f);
const float* orc_filter_spectra(const orc_filter* f); /* [P][B+1] interleaved complex */

typedef struct orc_blockconv orc_blockconv;
orc_blockconv* orc_blockconv_create(unsigned block, unsigned max_partitions);
void orc_blockconv_destroy(orc_blockconv* bc);
void orc_blockconv_set_filter(orc_blockconv* bc, const orc_filter* f, int crossfade);
void orc_blockconv_convolve(orc_blockconv* bc, const float* in, float* out);

/* ---- convolver.c : SURVEY.md 8.A multichannel Convolver ---- */
enum { ORC_MODE_PER_CHANNEL = 0, ORC_MODE_ROUTED = 1, ORC_MODE_MIMO = 2 };
typedef struct orc_convolver orc_convolver;
orc_convolver* orc_convolver_create(unsigned block, unsigned max_partitions, unsigned n_inputs, unsigned n_outputs,
                                    unsigned n_paths, int mode, unsigned ring_len, int fractional_delay,
                                    int nthreads);
void orc_convolver_destroy(orc_convolver* c);
void orc_convolver_set_route(orc_convolver* c, unsigned path, unsigned input, unsigned output, float gain);
void orc_convolver_set_filter(orc_convolver* c, unsigned path, const orc_filter* f, int crossfade, double delay);
/* nframes must be a multiple of block */
int orc_convolver_process(orc_convolver* c, const void* in, int infmt, int in_be, unsigned in_channels, void* out,
                          int outfmt, int out_be, unsigned out_channels, unsigned nframes);

/* ---- direct.c : float64 truth ---- */
/* y[n] = sum_j h[j] x[n-j] for n in [n0, n0+count), zero history before x[0] */
void orc_direct_convolve(const double* x, unsigned nx, const double* h, unsigned nh, unsigned n0, unsigned count,
                         double* y, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
