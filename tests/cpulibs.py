"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE).

  oracle/liboracle.so        plain-C restatement (prefix ``orc_``), always available
  oracle/_ref/libbbcref.so   the reference's own in-tree code (prefix ``ref_``); prebuilt in the
                             dev container, travels to the GPU box as a built file

Both expose the same signatures for the pieces that exist in the reference tree, so a
``CpuLib`` wraps either.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

FMT_16, FMT_24, FMT_32, FMT_FLOAT, FMT_DOUBLE = 1, 2, 3, 4, 5
FMT_BYTES = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8}
FMT_NAMES = {1: "s16", 2: "s24", 3: "s32", 4: "f32", 5: "f64"}

u32 = C.c_uint
vp = C.c_void_p


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(vp)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])


class CpuLib:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        p = prefix

        def fn(name, res, args):
            f = getattr(L, p + name)
            f.restype = res
            f.argtypes = args
            return f

        self._bits = fn("get_bits_per_sample", u32, [C.c_int])
        self._bytes = fn("get_bytes_per_sample", u32, [C.c_int])
        self._sanity = fn("block_transfer_sanity_checks", C.c_int, [C.POINTER(u32)] * 6 + [C.c_int])
        self._transfer = fn("transfer_samples", None,
                            [vp, C.c_int, C.c_int, u32, u32, vp, C.c_int, C.c_int, u32, u32, u32, u32])
        self._linear = fn("transfer_samples_linear", None, [vp, C.c_int, vp, C.c_int, u32])
        self._ditherer = fn("transfer_samples_ditherer", u32,
                            [vp, C.c_int, C.c_int, u32, u32, vp, C.c_int, C.c_int, u32, u32, u32, u32, C.c_int])
        if prefix == "orc_":  # libbbx's own TPDF law and call-site table: restated in the oracle only
            self._tpdf = fn("transfer_samples_dither", None,
                            [vp, C.c_int, C.c_int, u32, u32, vp, C.c_int, C.c_int, u32, u32, u32, u32, C.c_int, C.c_uint64, vp, vp])
            self._dbits = fn("dither_bits", C.c_int, [C.c_int, C.c_int])
        self._mix32 = fn("mix_samples_f32", None, [vp, u32, u32, vp, u32, u32, u32, u32, C.c_float])
        self._mix64 = fn("mix_samples_f64", None, [vp, u32, u32, vp, u32, u32, u32, u32, C.c_double])
        self._mixi = fn("mix_samples_interp", None, [vp, u32, u32, vp, u32, u32, u32, u32, vp, C.c_float])
        self._istep = fn("interpolator_step", None, [vp, C.c_float, u32])
        self._fadd = fn("fractional_sample_additional_delay_required", u32, [])
        self._frac32 = fn("fractional_samples_f32", None, [vp, u32, u32, u32, vp, u32, vp])
        self._frac64 = fn("fractional_samples_f64", None, [vp, u32, u32, u32, vp, u32, vp])
        self._dl_create = fn("delay_create", vp, [])
        self._dl_destroy = fn("delay_destroy", None, [vp])
        self._dl_set_size = fn("delay_set_size", None, [vp, u32, u32, C.c_int])
        self._dl_channels = fn("delay_get_channels", u32, [vp])
        self._dl_length = fn("delay_get_length", u32, [vp])
        self._dl_wpos = fn("delay_get_write_position", u32, [vp])
        self._dl_format = fn("delay_get_format", C.c_int, [vp])
        self._dl_write = fn("delay_write_samples", u32, [vp, vp, C.c_int, u32, u32, u32])
        self._dl_inc = fn("delay_increment_write_position", None, [vp, u32])
        self._dl_read = fn("delay_read_samples", u32, [vp, vp, C.c_int, u32, u32, u32, u32])
        self._dl_copy = fn("delay_copy_buffer", u32, [vp, vp, u32])
        self._rg_create = fn("ring_create", vp, [])
        self._rg_rpos = fn("ring_get_read_position", u32, [vp])
        self._rg_ravail = fn("ring_get_read_frames_available", u32, [vp])
        self._rg_wavail = fn("ring_get_write_frames_available", u32, [vp])
        self._rg_rinc = fn("ring_increment_read_position", None, [vp, u32])
        self._mlb_create = fn("mlb_create", vp, [u32, u32])
        self._mlb_destroy = fn("mlb_destroy", None, [vp])
        self._mlb_write = fn("mlb_write_layer", None, [vp, u32, vp, u32, u32, u32, u32, u32])
        self._mlb_avail = fn("mlb_available_frames", u32, [vp])
        self._mlb_read = fn("mlb_read_buffer", u32, [vp, u32, vp, u32, u32, u32, u32, C.c_int])
        dbl = C.c_double
        self._bq_coeffs = fn("biquad_calc_coeffs", None, [C.c_int, dbl, dbl, dbl, dbl, vp])
        self._bq_create = fn("biquad_create", vp, [u32])
        self._bq_destroy = fn("biquad_destroy", None, [vp])
        self._bq_set = fn("biquad_set_coeffs", None, [vp, vp, dbl])
        self._bq_calc = fn("biquad_calc", None, [vp, C.c_int, dbl, dbl, dbl, dbl, dbl])
        self._bq_process = fn("biquad_process", None, [vp, vp, vp, u32, u32, u32, u32])
        self._bq_state = fn("biquad_get_state", None, [vp, vp, vp, vp])
        self._bq_reset = fn("biquad_reset", None, [vp])
        self._fb_create = fn("fbank_create", vp, [u32, u32])
        self._fb_destroy = fn("fbank_destroy", None, [vp])
        self._fb_set_filters = fn("fbank_set_filters", None, [vp, u32])
        self._fb_add = fn("fbank_add_filter", None, [vp, vp])
        self._fb_set_channels = fn("fbank_set_channels", None, [vp, u32])
        self._fb_set = fn("fbank_set_coeffs", None, [vp, u32, vp, dbl])
        self._fb_calc = fn("fbank_calc", None, [vp, u32, C.c_int, dbl, dbl, dbl, dbl, dbl])
        self._fb_process = fn("fbank_process", None, [vp, vp, vp, u32, u32, u32, u32])
        self._fb_state = fn("fbank_get_state", None, [vp, u32, vp, vp, vp])
        self._fb_reset = fn("fbank_reset", None, [vp])
        self._ap_create = fn("allpass_create", vp, [u32, u32, vp, vp])
        self._ap_destroy = fn("allpass_destroy", None, [vp])
        self._ap_process = fn("allpass_process", None, [vp, vp, vp, u32, u32, u32, u32, u32])
        self._ap_state = fn("allpass_get_state", u32, [vp, u32, vp, u32])
        lng = C.c_long
        self._cs_create = fn("cascade_create", vp, [u32, u32, C.c_int, C.c_int])
        self._cs_destroy = fn("cascade_destroy", None, [vp])
        self._cs_set = fn("cascade_set_coefficients", C.c_int, [vp, u32, vp, u32])
        self._cs_reset = fn("cascade_reset", None, [vp])
        self._cs_process = fn("cascade_process", None, [vp, vp, lng, lng, vp, lng, lng, u32])
        self._cs_state = fn("cascade_get_state", u32, [vp, u32, vp, vp, vp, vp, vp])

    # ---- formats ----
    def bits_per_sample(self, fmt):
        return self._bits(fmt)

    def bytes_per_sample(self, fmt):
        return self._bytes(fmt)

    def sanity(self, src_channel, src_channels, dst_channel, dst_channels, nchannels, nframes, allow=True):
        v = [u32(x & 0xFFFFFFFF) for x in (src_channel, src_channels, dst_channel, dst_channels, nchannels, nframes)]
        ok = self._sanity(*[C.byref(x) for x in v], int(allow))
        return bool(ok), tuple(x.value for x in v)

    def transfer(self, src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel,
                 dst_channels, nchannels, nframes):
        """src/dst: uint8 numpy byte buffers (dst modified in place)."""
        self._transfer(_ptr(src), srctype, int(src_be), src_channel, src_channels, _ptr(dst), dsttype, int(dst_be),
                       dst_channel, dst_channels, nchannels & 0xFFFFFFFF, nframes)

    def transfer_linear(self, src, srctype, dst, dsttype, nsamples):
        self._linear(_ptr(src), srctype, _ptr(dst), dsttype, nsamples)

    def transfer_ditherer(self, src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel,
                          dst_channels, nchannels, nframes, mode):
        """TransferSamples with a Ditherer object: mode 0 = the no-op base class, 1 = the stateful test subclass of
        tests/cpp/test_ditherer.h.  Returns the number of Dither() calls the test subclass received."""
        return self._ditherer(_ptr(src), srctype, int(src_be), src_channel, src_channels, _ptr(dst), dsttype, int(dst_be),
                              dst_channel, dst_channels, nchannels & 0xFFFFFFFF, nframes, mode)

    def transfer_tpdf(self, src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel,
                      dst_channels, nchannels, nframes, seed):
        self._tpdf(_ptr(src), srctype, int(src_be), src_channel, src_channels, _ptr(dst), dsttype, int(dst_be),
                   dst_channel, dst_channels, nchannels & 0xFFFFFFFF, nframes, 1, seed, None, None)

    def dither_bits(self, srctype, dsttype):
        return self._dbits(srctype, dsttype)

    # ---- mixing ----
    def mix(self, src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul=1.0):
        if src.dtype == np.float32:
            self._mix32(_ptr(src), src_channel, src_channels, _ptr(dst), dst_channel, dst_channels,
                        nchannels & 0xFFFFFFFF, nframes, mul)
        else:
            self._mix64(_ptr(src), src_channel, src_channels, _ptr(dst), dst_channel, dst_channels,
                        nchannels & 0xFFFFFFFF, nframes, mul)

    def mix_interp(self, src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, state,
                   inc):
        """state: float32[2] = (target, current), advanced in place."""
        self._mixi(_ptr(src), src_channel, src_channels, _ptr(dst), dst_channel, dst_channels,
                   nchannels & 0xFFFFFFFF, nframes, _ptr(state), inc)

    def interp_step(self, state, inc, nsteps):
        self._istep(_ptr(state), inc, nsteps)

    # ---- fractional sample ----
    def frac_additional(self):
        return self._fadd()

    def frac(self, buffer, channel, channels, length, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        out = np.empty(pos.shape, dtype=np.float64)
        f = self._frac32 if buffer.dtype == np.float32 else self._frac64
        f(_ptr(buffer), channel, channels, length, _ptr(pos), pos.size, _ptr(out))
        return out

    # ---- delay buffer ----
    def delay(self, ring=False):
        return CpuDelay(self, ring)

    # ---- MultilayerBuffer<float> ----
    def multilayer(self, channels, layers):
        return CpuMultilayer(self, channels, layers)

    # ---- BiQuadCoeffs / BiQuad ----
    def biquad_coeffs(self, ftype, freq, fs, gain=0.0, bandwidth=1.0):
        out = np.zeros(5, dtype=np.float64)
        self._bq_coeffs(ftype, freq, fs, gain, bandwidth, _ptr(out))
        return out

    def biquad(self, channels):
        return CpuBiquad(self, channels)

    def fbank(self, channels, filters):
        return CpuFbank(self, channels, filters)

    def allpass(self, channels, delays, coeffs):
        return CpuAllpass(self, channels, delays, coeffs)

    def cascade(self, channels, numfilters, vectorise=True, unroll=True):
        return CpuCascade(self, channels, numfilters, vectorise, unroll)


class CpuCascade:
    """BiQuadCascade bank (one reference object / C restatement per channel)."""

    def __init__(self, lib, channels, numfilters, vectorise, unroll):
        self.l, self.channels = lib, channels
        self.h = lib._cs_create(channels, numfilters, int(vectorise), int(unroll))

    def close(self):
        if self.h:
            self.l._cs_destroy(self.h)
            self.h = None

    def set_coefficients(self, coeffs, channel=None):
        c = np.ascontiguousarray(coeffs, dtype=np.float32)
        return bool(self.l._cs_set(self.h, 0xFFFFFFFF if channel is None else channel, _ptr(c), c.size))

    def reset(self):
        self.l._cs_reset(self.h)

    def process(self, src, dst, nframes, interleaved=True):
        cs, fs = (1, self.channels) if interleaved else (nframes, 1)
        self.l._cs_process(self.h, _ptr(src), cs, fs, _ptr(dst), cs, fs, nframes)

    def state(self, channel):
        a = [np.zeros(12, dtype=np.float32) for _ in range(4)]
        last = np.zeros(1, dtype=np.float32)
        info = self.l._cs_state(self.h, channel, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(last))
        return a[0], a[1], a[2], a[3], last, info


class CpuAllpass:
    def __init__(self, lib, channels, delays, coeffs):
        self.l, self.channels, self.delays = lib, channels, list(delays)
        d = np.ascontiguousarray(delays, dtype=np.uint32)
        c = np.ascontiguousarray(coeffs, dtype=np.float32)
        self.h = lib._ap_create(channels, len(delays), _ptr(d), _ptr(c))

    def close(self):
        if self.h:
            self.l._ap_destroy(self.h)
            self.h = None

    def process(self, src, dst, srcchannel, nsrc, dstchannel, ndst, nframes):
        self.l._ap_process(self.h, _ptr(src), _ptr(dst), srcchannel, nsrc, dstchannel, ndst, nframes)

    def state(self, f):
        ring = np.zeros(self.channels * self.delays[f], dtype=np.float32)
        pos = self.l._ap_state(self.h, f, _ptr(ring), ring.size)
        return ring, pos


class CpuBiquad:
    def __init__(self, lib, channels):
        self.l, self.channels = lib, channels
        self.h = lib._bq_create(channels)

    def close(self):
        if self.h:
            self.l._bq_destroy(self.h)
            self.h = None

    def set_coeffs(self, c5, interp_samples=0.0):
        self.l._bq_set(self.h, _ptr(np.ascontiguousarray(c5, dtype=np.float64)), interp_samples)

    def calc(self, ftype, freq, fs, gain=0.0, bandwidth=1.0, interp_time=0.0):
        self.l._bq_calc(self.h, ftype, freq, fs, gain, bandwidth, interp_time)

    def process(self, src, dst, nchannels, nsrc, ndst, nframes):
        self.l._bq_process(self.h, _ptr(src), _ptr(dst), nchannels, nsrc, ndst, nframes)

    def state(self):
        w = np.zeros(2 * max(1, self.channels), dtype=np.float64)
        cur, md = np.zeros(5, dtype=np.float64), np.zeros(2, dtype=np.float64)
        self.l._bq_state(self.h, _ptr(w), _ptr(cur), _ptr(md))
        return w[:2 * self.channels], cur, md

    def reset(self):
        self.l._bq_reset(self.h)


class CpuFbank:
    """BiQuadFilterBank: the reference's own class (oracle/_ref) or the C restatement of its filter-by-filter loop."""

    def __init__(self, lib, channels, filters):
        self.l, self.channels, self.filters = lib, channels, filters
        self.h = lib._fb_create(channels, filters)

    def close(self):
        if self.h:
            self.l._fb_destroy(self.h)
            self.h = None

    def set_filters(self, n):
        self.l._fb_set_filters(self.h, n)
        self.filters = n

    def add_filter(self, c5):
        self.l._fb_add(self.h, _ptr(np.ascontiguousarray(c5, dtype=np.float64)))
        self.filters += 1

    def set_channels(self, n):
        self.l._fb_set_channels(self.h, n)
        self.channels = n

    def set_coeffs(self, filter, c5, interp_samples=0.0):
        self.l._fb_set(self.h, filter, _ptr(np.ascontiguousarray(c5, dtype=np.float64)), interp_samples)

    def calc(self, filter, ftype, freq, fs, gain=0.0, bandwidth=1.0, interp_time=0.0):
        self.l._fb_calc(self.h, filter, ftype, freq, fs, gain, bandwidth, interp_time)

    def process(self, src, dst, nchannels, nsrc, ndst, nframes):
        self.l._fb_process(self.h, _ptr(src), _ptr(dst), nchannels, nsrc, ndst, nframes)

    def state(self, filter):
        w = np.zeros(2 * max(1, self.channels), dtype=np.float64)
        cur, md = np.zeros(5, dtype=np.float64), np.zeros(2, dtype=np.float64)
        self.l._fb_state(self.h, filter, _ptr(w), _ptr(cur), _ptr(md))
        return w[:2 * self.channels], cur, md

    def reset(self):
        self.l._fb_reset(self.h)


class CpuMultilayer:
    def __init__(self, lib, channels, layers):
        self.l = lib
        self.h = lib._mlb_create(channels, layers)

    def close(self):
        if self.h:
            self.l._mlb_destroy(self.h)
            self.h = None

    def write_layer(self, layer, src, srcchannel, nsrcchannels, dstchannel, nchannels, nframes):
        self.l._mlb_write(self.h, layer, _ptr(src), srcchannel, nsrcchannels, dstchannel, nchannels & 0xFFFFFFFF, nframes)

    def available(self):
        return self.l._mlb_avail(self.h)

    def read(self, srcchannel, dst, dstchannel, ndstchannels, nchannels, nframes, overwrite=True):
        return self.l._mlb_read(self.h, srcchannel, _ptr(dst), dstchannel, ndstchannels, nchannels & 0xFFFFFFFF, nframes,
                                int(overwrite))


class CpuDelay:
    def __init__(self, lib, ring=False):
        self.l = lib
        self.h = lib._rg_create() if ring else lib._dl_create()

    read_position = property(lambda s: s.l._rg_rpos(s.h))
    read_available = property(lambda s: s.l._rg_ravail(s.h))
    write_available = property(lambda s: s.l._rg_wavail(s.h))

    def increment_read(self, nframes):
        self.l._rg_rinc(self.h, nframes)

    def close(self):
        if self.h:
            self.l._dl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_size(self, chans, length, fmt=FMT_FLOAT):
        self.l._dl_set_size(self.h, chans, length, fmt)

    channels = property(lambda s: s.l._dl_channels(s.h))
    length = property(lambda s: s.l._dl_length(s.h))
    write_position = property(lambda s: s.l._dl_wpos(s.h))
    format = property(lambda s: s.l._dl_format(s.h))

    def write(self, src, srcformat, channel, nchannels, nframes):
        return self.l._dl_write(self.h, _ptr(src), srcformat, channel, nchannels & 0xFFFFFFFF, nframes)

    def increment(self, nframes):
        self.l._dl_inc(self.h, nframes)

    def read(self, dst, dstformat, delay, channel, nchannels, nframes):
        return self.l._dl_read(self.h, _ptr(dst), dstformat, delay, channel, nchannels & 0xFFFFFFFF, nframes)

    def raw(self):
        n = self.channels * self.length * FMT_BYTES[self.format]
        out = np.zeros(n, dtype=np.uint8)
        got = self.l._dl_copy(self.h, _ptr(out), n)
        assert got == n
        return out


class Oracle(CpuLib):
    """liboracle.so: adds FFT / UPOLS / Convolver / direct convolution."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        super().__init__(path, "orc_")
        L = self.lib
        L.orc_cfft.argtypes = [vp, u32, C.c_int]
        L.orc_rfft.argtypes = [vp, vp, u32]
        L.orc_irfft.argtypes = [vp, vp, u32]
        L.orc_filter_create.restype = vp
        L.orc_filter_create.argtypes = [vp, u32, u32]
        L.orc_filter_destroy.argtypes = [vp]
        L.orc_filter_partitions.restype = u32
        L.orc_filter_partitions.argtypes = [vp]
        L.orc_filter_spectra.restype = C.POINTER(C.c_float)
        L.orc_filter_spectra.argtypes = [vp]
        L.orc_blockconv_create.restype = vp
        L.orc_blockconv_create.argtypes = [u32, u32]
        L.orc_blockconv_destroy.argtypes = [vp]
        L.orc_blockconv_set_filter.argtypes = [vp, vp, C.c_int]
        L.orc_blockconv_convolve.argtypes = [vp, vp, vp]
        L.orc_convolver_create.restype = vp
        L.orc_convolver_create.argtypes = [u32, u32, u32, u32, u32, C.c_int, u32, C.c_int, C.c_int]
        L.orc_convolver_destroy.argtypes = [vp]
        L.orc_convolver_set_route.argtypes = [vp, u32, u32, u32, C.c_float]
        L.orc_convolver_set_filter.argtypes = [vp, u32, vp, C.c_int, C.c_double]
        L.orc_convolver_process.restype = C.c_int
        L.orc_convolver_process.argtypes = [vp, vp, C.c_int, C.c_int, u32, vp, C.c_int, C.c_int, u32, u32]
        L.orc_direct_convolve.argtypes = [vp, u32, vp, u32, u32, u32, vp, C.c_int]

    def rfft(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(x.size + 2, dtype=np.float32)
        self.lib.orc_rfft(_ptr(x), _ptr(out), x.size)
        return out.view(np.complex64)

    def irfft(self, X, n):
        X = np.ascontiguousarray(X, dtype=np.complex64)
        out = np.zeros(n, dtype=np.float32)
        self.lib.orc_irfft(_ptr(X.view(np.float32)), _ptr(out), n)
        return out

    def cfft(self, z, inverse=False):
        z = np.array(z, dtype=np.complex64)
        self.lib.orc_cfft(_ptr(z.view(np.float32)), z.size, int(inverse))
        return z

    def filter(self, ir, block):
        return OracleFilter(self, ir, block)

    def blockconv(self, block, max_partitions):
        return OracleBlockConv(self, block, max_partitions)

    def convolver(self, **kw):
        return OracleConvolver(self, **kw)

    def direct(self, x, h, n0=0, count=None, nthreads=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        h = np.ascontiguousarray(h, dtype=np.float64)
        if count is None:
            count = x.size - n0
        y = np.zeros(count, dtype=np.float64)
        self.lib.orc_direct_convolve(_ptr(x), x.size, _ptr(h), h.size, n0, count, _ptr(y),
                                     nthreads or (os.cpu_count() or 1))
        return y


class OracleFilter:
    def __init__(self, orc, ir, block):
        self.orc = orc
        ir = np.ascontiguousarray(ir, dtype=np.float32)
        self.h = orc.lib.orc_filter_create(_ptr(ir), ir.size, block)
        self.block = block
        self.partitions = orc.lib.orc_filter_partitions(self.h)

    def spectra(self):
        K = self.block + 1
        p = self.orc.lib.orc_filter_spectra(self.h)
        a = np.ctypeslib.as_array(p, shape=(self.partitions * K * 2,)).copy()
        return a.view(np.complex64).reshape(self.partitions, K)

    def __del__(self):
        try:
            if self.h:
                self.orc.lib.orc_filter_destroy(self.h)
                self.h = None
        except Exception:
            pass


class OracleBlockConv:
    def __init__(self, orc, block, max_partitions):
        self.orc = orc
        self.block = block
        self.h = orc.lib.orc_blockconv_create(block, max_partitions)
        self._keep = []

    def set_filter(self, f, crossfade=False):
        self._keep.append(f)
        self.orc.lib.orc_blockconv_set_filter(self.h, f.h if f is not None else None, int(crossfade))

    def convolve(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.size == self.block
        y = np.zeros(self.block, dtype=np.float32)
        self.orc.lib.orc_blockconv_convolve(self.h, _ptr(x), _ptr(y))
        return y

    def __del__(self):
        try:
            if self.h:
                self.orc.lib.orc_blockconv_destroy(self.h)
                self.h = None
        except Exception:
            pass


MODE_PER_CHANNEL, MODE_ROUTED, MODE_MIMO = 0, 1, 2


class OracleConvolver:
    def __init__(self, orc, block, max_partitions, n_inputs, n_outputs=None, n_paths=None, mode=MODE_PER_CHANNEL,
                 ring_len=0, fractional_delay=False, nthreads=None):
        self.orc = orc
        self.block = block
        n_outputs = n_inputs if n_outputs is None else n_outputs
        if mode == MODE_PER_CHANNEL:
            n_paths = n_inputs
        elif mode == MODE_MIMO:
            n_paths = n_inputs * n_outputs
        self.n_inputs, self.n_outputs, self.n_paths = n_inputs, n_outputs, n_paths
        self.h = orc.lib.orc_convolver_create(block, max_partitions, n_inputs, n_outputs, n_paths, mode, ring_len,
                                              int(fractional_delay), nthreads or (os.cpu_count() or 1))
        self._keep = []

    def set_route(self, path, inp, out, gain=1.0):
        self.orc.lib.orc_convolver_set_route(self.h, path, inp, out, gain)

    def set_filter(self, path, f, crossfade=False, delay=0.0):
        self._keep.append(f)
        self.orc.lib.orc_convolver_set_filter(self.h, path, f.h if f is not None else None, int(crossfade), delay)

    def process(self, inp, infmt, in_channels, outfmt, out_channels, nframes, in_be=False, out_be=False, out=None):
        """inp: uint8 byte buffer (or typed array) of interleaved PCM; returns uint8 byte buffer."""
        inp = np.ascontiguousarray(inp).view(np.uint8).reshape(-1)
        assert inp.size == nframes * in_channels * FMT_BYTES[infmt]
        if out is None:
            out = np.zeros(nframes * out_channels * FMT_BYTES[outfmt], dtype=np.uint8)
        rc = self.orc.lib.orc_convolver_process(self.h, _ptr(inp), infmt, int(in_be), in_channels, _ptr(out), outfmt,
                                                int(out_be), out_channels, nframes)
        assert rc == 0, rc
        return out

    def __del__(self):
        try:
            if self.h:
                self.orc.lib.orc_convolver_destroy(self.h)
                self.h = None
        except Exception:
            pass


_ORACLE = None
_REF = None


def oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def ref_path():
    return os.path.join(ORACLE_DIR, "_ref", "libbbcref.so")


def have_ref():
    return os.path.exists(ref_path())


def reference():
    """The reference's own code (oracle/_ref); None when it was never built."""
    global _REF
    if _REF is None and have_ref():
        _REF = CpuLib(ref_path(), "ref_")
    return _REF
