"""bbcat-dsp_b200 -- Python host binding of libbbx.so (the B200 partitioned-convolution engine).

Thin ctypes layer over the C ABI in include/bbx.h; names follow the bbcat-dsp interfaces the ABI
replaces (TransferSamples, MixSamples, FractionalSample, SoundDelayBuffer, BlockConvolver, Convolver).
There is no CPU path here: importing works without a GPU (so the symbol table can be checked), but
every compute call raises BbxError when the CUDA library or device is missing.

The directory name carries a hyphen; import it as ``bbcat_dsp_b200`` (see bbcat_dsp_b200.py at the
repo root).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BBX_LIB: another build of the same library (A/B runs of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("BBX_LIB") or os.path.join(_HERE, "libbbx.so")

FMT_UNKNOWN, FMT_16BIT, FMT_24BIT, FMT_32BIT, FMT_FLOAT, FMT_DOUBLE = 0, 1, 2, 3, 4, 5
FMT_BYTES = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8}
MODE_PER_CHANNEL, MODE_ROUTED, MODE_MIMO = 0, 1, 2

u8, u32, u64, vp = C.c_uint8, C.c_uint32, C.c_uint64, C.c_void_p


class BbxError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("block_size", u32), ("max_partitions", u32), ("n_inputs", u32),
                ("n_outputs", u32), ("n_paths", u32), ("mode", C.c_int), ("max_blocks", u32), ("max_delay", u32),
                ("fractional_delay", C.c_int), ("ring_length", u32), ("mac_ctas_per_sm", u32), ("mac_l2_keep_16ths", u32),
                ("mac_time_tile", u32), ("mimo_tensor", u32), ("mimo_shard_world", u32), ("mimo_shard_rank", u32), ("reserved", u32 * 2)]


# every symbol include/bbx.h declares: name -> (restype, argtypes)
_RECT = [u32, u32]
SYMBOLS = {
    "bbx_version": (C.c_int, []),
    "bbx_last_error": (C.c_char_p, []),
    "bbx_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "bbx_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "bbx_host_free": (C.c_int, [vp]),
    "bbx_shard_range": (C.c_int, [u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]),
    "bbx_get_bits_per_sample": (u8, [C.c_int]),
    "bbx_get_bytes_per_sample": (u8, [C.c_int]),
    "bbx_block_transfer_sanity_checks": (C.c_int, [C.POINTER(u32)] * 6 + [C.c_int]),
    "bbx_transfer_samples": (C.c_int, [vp, C.c_int, C.c_int, u32, u32, vp, C.c_int, C.c_int, u32, u32, u32, u32]),
    "bbx_transfer_samples_dev": (C.c_int, [vp, C.c_int, C.c_int, u32, u32, vp, C.c_int, C.c_int, u32, u32, u32, u32, vp]),
    "bbx_transfer_samples_linear": (C.c_int, [vp, C.c_int, vp, C.c_int, u32]),
    "bbx_dither_bits": (C.c_int, [C.c_int, C.c_int]),
    "bbx_transfer_samples_dither": (C.c_int, [vp, C.c_int, C.c_int, u32, u32, vp, C.c_int, C.c_int, u32, u32, u32, u32, C.c_int, u64]),
    "bbx_transfer_samples_dither_dev": (C.c_int, [vp, C.c_int, C.c_int, u32, u32, vp, C.c_int, C.c_int, u32, u32, u32, u32, C.c_int, u64, vp]),
    "bbx_mix_samples_f32": (C.c_int, [vp, u32, u32, vp, u32, u32, u32, u32, C.c_float]),
    "bbx_mix_samples_f64": (C.c_int, [vp, u32, u32, vp, u32, u32, u32, u32, C.c_double]),
    "bbx_mix_samples_interp": (C.c_int, [vp, u32, u32, vp, u32, u32, u32, u32, vp, C.c_float]),
    "bbx_mix_samples_f32_dev": (C.c_int, [vp, u32, u32, vp, u32, u32, u32, u32, C.c_float, vp]),
    "bbx_mix_samples_interp_dev": (C.c_int, [vp, u32, u32, vp, u32, u32, u32, u32, vp, C.c_float, vp]),
    "bbx_interpolator_step": (C.c_int, [vp, C.c_float, u32]),
    "bbx_fractional_sample_additional_delay_required": (u32, []),
    "bbx_fractional_samples_f32": (C.c_int, [vp, u32, u32, u32, vp, u32, vp]),
    "bbx_fractional_samples_f64": (C.c_int, [vp, u32, u32, u32, vp, u32, vp]),
    "bbx_fractional_samples_f32_dev": (C.c_int, [vp, u32, u32, u32, vp, u32, vp, vp]),
    "bbx_delay_create": (C.c_int, [C.POINTER(vp)]),
    "bbx_delay_destroy": (C.c_int, [vp]),
    "bbx_delay_set_size": (C.c_int, [vp, u32, u32, C.c_int]),
    "bbx_delay_get_channels": (u32, [vp]),
    "bbx_delay_get_length": (u32, [vp]),
    "bbx_delay_get_write_position": (u32, [vp]),
    "bbx_delay_get_format": (C.c_int, [vp]),
    "bbx_delay_write_samples": (u32, [vp, vp, C.c_int, u32, u32, u32]),
    "bbx_delay_increment_write_position": (C.c_int, [vp, u32]),
    "bbx_delay_read_samples": (u32, [vp, vp, C.c_int, u32, u32, u32, u32]),
    "bbx_delay_read_sample": (C.c_float, [vp, u32, u32]),
    "bbx_delay_get_buffer_dev": (vp, [vp]),
    "bbx_delay_copy_buffer": (u32, [vp, vp, u32]),
    "bbx_ring_create": (C.c_int, [C.POINTER(vp)]),
    "bbx_ring_get_read_position": (u32, [vp]),
    "bbx_ring_get_read_frames_available": (u32, [vp]),
    "bbx_ring_get_write_frames_available": (u32, [vp]),
    "bbx_ring_increment_read_position": (C.c_int, [vp, u32]),
    "bbx_mlb_create": (C.c_int, [u32, u32, C.POINTER(vp)]),
    "bbx_mlb_destroy": (C.c_int, [vp]),
    "bbx_mlb_get_channels": (u32, [vp]),
    "bbx_mlb_get_layers": (u32, [vp]),
    "bbx_mlb_get_available_frames": (u32, [vp]),
    "bbx_mlb_write_layer": (C.c_int, [vp, u32, vp, u32, u32, u32, u32, u32]),
    "bbx_mlb_read_buffer": (u32, [vp, u32, vp, u32, u32, u32, u32, C.c_int]),
    "bbx_engine_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
    "bbx_engine_destroy": (C.c_int, [vp]),
    "bbx_engine_get_ring_length": (u32, [vp]),
    "bbx_engine_get_stream": (vp, [vp]),
    "bbx_filter_create": (C.c_int, [vp, vp, u32, C.POINTER(vp)]),
    "bbx_filter_destroy": (C.c_int, [vp]),
    "bbx_filter_partitions": (u32, [vp]),
    "bbx_filter_read_spectra": (C.c_int, [vp, vp, C.c_size_t]),
    "bbx_set_route": (C.c_int, [vp, u32, u32, u32, C.c_float]),
    "bbx_set_filter": (C.c_int, [vp, u32, vp, C.c_int, C.c_double]),
    "bbx_set_filters": (C.c_int, [vp, u32, vp, vp, vp, vp]),
    "bbx_process": (C.c_int, [vp, vp, C.c_int, C.c_int, u32, vp, C.c_int, C.c_int, u32, u32]),
    "bbx_process_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, u32, vp, C.c_int, C.c_int, u32, u32]),
    "bbx_process_async": (C.c_int, [vp, vp, C.c_int, C.c_int, u32, vp, C.c_int, C.c_int, u32, u32]),
    "bbx_engine_sync": (C.c_int, [vp]),
    "bbx_blockconvolver_convolve": (C.c_int, [vp, vp, vp]),
    "bbx_engine_timer_start": (C.c_int, [vp]),
    "bbx_engine_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "bbx_engine_launch_count": (u64, [vp]),
    "bbx_engine_mac_kernel": (C.c_char_p, [vp]),
    "bbx_engine_profile_mac": (C.c_int, [vp, C.c_int]),
    "bbx_engine_mac_time": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
    "bbx_engine_exchange_time": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(u64), C.POINTER(u64)]),
    "bbx_engine_state_size": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
    "bbx_engine_get_state": (C.c_int, [vp, vp, C.c_size_t]),
    "bbx_engine_set_state": (C.c_int, [vp, vp, C.c_size_t]),
    "bbx_engine_io_trace": (C.c_int, [vp, u32]),
    "bbx_engine_io_trace_read": (C.c_int, [vp, C.POINTER(C.c_float), u32, C.POINTER(u32)]),
    "bbx_engine_set_tuning": (C.c_int, [vp, u32, u32, u32]),
    "bbx_engine_flush_l2": (C.c_int, [vp, C.c_size_t]),
    "bbx_engine_peer_export": (C.c_int, [vp, vp]),
    "bbx_engine_peer_attach": (C.c_int, [vp, vp]),
    "bbx_engine_set_direct_io": (C.c_int, [vp, C.c_size_t]),
    "bbx_engine_set_mixdown_kernel": (C.c_int, [vp, C.c_int]),
    "bbx_engine_direct_calls": (u64, [vp]),
    "bbx_engine_set_fused": (C.c_int, [vp, C.c_int, u32]),
    "bbx_engine_fused_calls": (u64, [vp]),
    "bbx_biquad_calc_coeffs": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "bbx_biquad_create": (C.c_int, [u32, C.POINTER(vp)]),
    "bbx_biquad_destroy": (C.c_int, [vp]),
    "bbx_biquad_set_coeffs": (C.c_int, [vp, C.POINTER(C.c_double), C.c_double]),
    "bbx_biquad_calc": (C.c_int, [vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "bbx_biquad_process": (C.c_int, [vp, vp, vp, u32, u32, u32, u32]),
    "bbx_biquad_process_dev": (C.c_int, [vp, vp, vp, u32, u32, u32, u32, vp]),
    "bbx_biquad_get_state": (C.c_int, [vp, vp, vp, vp]),
    "bbx_biquad_reset": (C.c_int, [vp]),
    "bbx_fbank_create": (C.c_int, [u32, u32, C.POINTER(vp)]),
    "bbx_fbank_destroy": (C.c_int, [vp]),
    "bbx_fbank_set_filters": (C.c_int, [vp, u32]),
    "bbx_fbank_add_filter": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "bbx_fbank_set_channels": (C.c_int, [vp, u32]),
    "bbx_fbank_get_size": (C.c_int, [vp, C.POINTER(u32), C.POINTER(u32)]),
    "bbx_fbank_set_coeffs": (C.c_int, [vp, u32, C.POINTER(C.c_double), C.c_double]),
    "bbx_fbank_calc": (C.c_int, [vp, u32, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "bbx_fbank_process": (C.c_int, [vp, vp, vp, u32, u32, u32, u32]),
    "bbx_fbank_process_dev": (C.c_int, [vp, vp, vp, u32, u32, u32, u32, vp]),
    "bbx_fbank_get_state": (C.c_int, [vp, u32, vp, vp, vp]),
    "bbx_fbank_reset": (C.c_int, [vp]),
    "bbx_fbank_launches": (C.c_int, [vp, C.POINTER(u64)]),
    "bbx_probe_launch_sync": (C.c_int, [C.c_int, u32, u32, C.POINTER(C.c_double)]),
    "bbx_block_latency": (C.c_int, [vp, vp, C.c_int, C.c_int, u32, vp, C.c_int, C.c_int, u32, u32, u32, u32, C.POINTER(C.c_double)]),
    "bbx_sofa_open": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "bbx_sofa_open_memory": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "bbx_sofa_close": (C.c_int, [vp]),
    "bbx_sofa_get_sizes": (C.c_int, [vp] + [C.POINTER(u32)] * 4),
    "bbx_sofa_get_samplerate": (C.c_int, [vp, u32, C.POINTER(C.c_double)]),
    "bbx_sofa_get_ir": (C.c_int, [vp, u32, u32, u32, vp, u32]),
    "bbx_sofa_get_delay": (C.c_int, [vp, u32, u32, u32, C.POINTER(C.c_double)]),
    "bbx_sofa_get_position": (C.c_int, [vp, C.c_int, u32, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "bbx_sofa_nearest_measurement": (C.c_int, [vp, C.POINTER(C.c_double), C.c_int, C.POINTER(u32)]),
    "bbx_sofa_get_attribute": (C.c_int, [vp, C.c_char_p, C.c_char_p, u32]),
    "bbx_sofa_create_filters": (C.c_int, [vp, vp, u32, u32, C.POINTER(vp), u32]),
    "bbx_allpass_create": (C.c_int, [u32, u32, C.POINTER(u32), C.POINTER(C.c_float), C.POINTER(vp)]),
    "bbx_allpass_destroy": (C.c_int, [vp]),
    "bbx_allpass_process": (C.c_int, [vp, vp, vp, u32, u32, u32, u32, u32]),
    "bbx_allpass_process_dev": (C.c_int, [vp, vp, vp, u32, u32, u32, u32, u32, vp]),
    "bbx_allpass_get_state": (u32, [vp, u32, vp, u32]),
    "bbx_cascade_create": (C.c_int, [u32, u32, C.c_int, C.c_int, C.POINTER(vp)]),
    "bbx_cascade_destroy": (C.c_int, [vp]),
    "bbx_cascade_set_coefficients": (C.c_int, [vp, u32, vp, u32]),
    "bbx_cascade_reset": (C.c_int, [vp]),
    "bbx_cascade_process": (C.c_int, [vp, vp, vp, u32, C.c_int]),
    "bbx_cascade_process_dev": (C.c_int, [vp, vp, C.c_longlong, C.c_longlong, vp, C.c_longlong, C.c_longlong, u32, vp]),
    "bbx_cascade_get_state": (u32, [vp, u32, vp, vp, vp, vp, vp]),
    "bbx_engine_tensor_status": (C.c_int, [vp, C.POINTER(u64), C.POINTER(C.c_int)]),
    "bbx_engine_tensor_trace": (C.c_int, [vp, vp, u32]),
    "bbx_probe_fp32_tflops": (C.c_int, [C.c_int, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "bbx_comm_available": (C.c_int, []),
    "bbx_comm_unique_id": (C.c_int, [C.POINTER(u8)]),
    "bbx_comm_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(u8), C.c_int, C.POINTER(vp)]),
    "bbx_comm_destroy": (C.c_int, [vp]),
    "bbx_engine_set_comm": (C.c_int, [vp, vp]),
}

_lib = None


def lib():
    """Load libbbx.so (once).  Fails loudly: there is no other implementation to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BbxError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C bbcat-dsp_b200/csrc`; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise BbxError("libbbx error %d: %s" % (rc, lib().bbx_last_error().decode("utf-8", "replace")))


def _p(a):
    """void* of a numpy array / integer device pointer / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return vp(a)
    assert a.flags["C_CONTIGUOUS"], "buffer must be C-contiguous"
    return a.ctypes.data_as(vp)


def device_count():
    n = C.c_int(0)
    rc = lib().bbx_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def shard_range(nchannels, rank, world):
    first, count = u32(0), u32(0)
    _check(lib().bbx_shard_range(nchannels, rank, world, C.byref(first), C.byref(count)))
    return first.value, count.value


# ---- a1/a2 --------------------------------------------------------------------------------
def GetBitsPerSample(fmt):
    return lib().bbx_get_bits_per_sample(fmt)


def GetBytesPerSample(fmt):
    return lib().bbx_get_bytes_per_sample(fmt)


def BlockTransferSanityChecks(src_channel, src_channels, dst_channel, dst_channels, nchannels, nframes,
                              allowsinglechannel=True):
    v = [u32(x & 0xFFFFFFFF) for x in (src_channel, src_channels, dst_channel, dst_channels, nchannels, nframes)]
    ok = lib().bbx_block_transfer_sanity_checks(*[C.byref(x) for x in v], int(allowsinglechannel))
    return bool(ok), tuple(x.value for x in v)


# ---- a3-a7 --------------------------------------------------------------------------------
DITHER_NONE, DITHER_TPDF = 0, 1


def DitherBits(srctype, dsttype):
    """bit count the reference hands to Ditherer::Dither for this converter, -1 when it has no dither call site."""
    return lib().bbx_dither_bits(srctype, dsttype)


def TransferSamples(src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel, dst_channels,
                    nchannels=0xFFFFFFFF, nframes=1, dither=DITHER_NONE, seed=0):
    """Host byte buffers (numpy uint8 or typed arrays); dst is modified in place.  dither=DITHER_TPDF: the device-side
    TPDF ditherer (a7)."""
    _check(lib().bbx_transfer_samples_dither(_p(src), srctype, int(src_be), src_channel, src_channels, _p(dst), dsttype,
                                             int(dst_be), dst_channel, dst_channels, nchannels & 0xFFFFFFFF, nframes,
                                             dither, seed))


def TransferSamplesDev(src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel,
                       dst_channels, nchannels, nframes, stream=None):
    """Device pointers (ints, e.g. torch.Tensor.data_ptr())."""
    _check(lib().bbx_transfer_samples_dev(_p(src), srctype, int(src_be), src_channel, src_channels, _p(dst), dsttype,
                                          int(dst_be), dst_channel, dst_channels, nchannels & 0xFFFFFFFF, nframes,
                                          _p(stream)))


def TransferSamplesLinear(src, srctype, dst, dsttype, nsamples=1):
    _check(lib().bbx_transfer_samples_linear(_p(src), srctype, _p(dst), dsttype, nsamples))


# ---- a8/a9 --------------------------------------------------------------------------------
def MixSamples(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul=1.0,
               interp=None, inc=0.0):
    """MixSamples<T> (float32/float64 arrays) or, with interp=float32[2] (target, current), the
    interpolated form; interp is advanced in place."""
    if interp is not None:
        _check(lib().bbx_mix_samples_interp(_p(src), src_channel, src_channels, _p(dst), dst_channel, dst_channels,
                                            nchannels & 0xFFFFFFFF, nframes, _p(interp), inc))
    elif src.dtype == np.float32:
        _check(lib().bbx_mix_samples_f32(_p(src), src_channel, src_channels, _p(dst), dst_channel, dst_channels,
                                         nchannels & 0xFFFFFFFF, nframes, mul))
    else:
        _check(lib().bbx_mix_samples_f64(_p(src), src_channel, src_channels, _p(dst), dst_channel, dst_channels,
                                         nchannels & 0xFFFFFFFF, nframes, mul))


def InterpolatorStep(state, inc, nsteps=1):
    _check(lib().bbx_interpolator_step(_p(state), inc, nsteps))


# ---- a10 ----------------------------------------------------------------------------------
def FractionalSampleAdditionalDelayRequired():
    return lib().bbx_fractional_sample_additional_delay_required()


def FractionalSample(buffer, channel, channels, length, pos):
    """Batched FractionalSample: pos may be a scalar or an array of positions."""
    scalar = np.isscalar(pos)
    pos = np.ascontiguousarray(np.atleast_1d(pos), dtype=np.float64)
    out = np.empty(pos.shape, dtype=np.float64)
    f = lib().bbx_fractional_samples_f32 if buffer.dtype == np.float32 else lib().bbx_fractional_samples_f64
    _check(f(_p(buffer), channel, channels, length, _p(pos), pos.size, _p(out)))
    return float(out[0]) if scalar else out


# ---- a11 ----------------------------------------------------------------------------------
class SoundDelayBuffer:
    def __init__(self):
        h = vp()
        _check(lib().bbx_delay_create(C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_delay_destroy(self.h)
            self.h = None

    __del__ = close

    def SetSize(self, chans, length, fmt=FMT_FLOAT):
        _check(lib().bbx_delay_set_size(self.h, chans, length, fmt))

    def GetChannels(self):
        return lib().bbx_delay_get_channels(self.h)

    def GetLength(self):
        return lib().bbx_delay_get_length(self.h)

    def GetWritePosition(self):
        return lib().bbx_delay_get_write_position(self.h)

    def GetFormat(self):
        return lib().bbx_delay_get_format(self.h)

    def WriteSamples(self, src, srcformat, channel=0, nchannels=0xFFFFFFFF, nframes=1):
        return lib().bbx_delay_write_samples(self.h, _p(src), srcformat, channel, nchannels & 0xFFFFFFFF, nframes)

    def IncrementWritePosition(self, nframes=1):
        _check(lib().bbx_delay_increment_write_position(self.h, nframes))

    def ReadSamples(self, dst, dstformat, delay, channel=0, nchannels=0xFFFFFFFF, nframes=1):
        return lib().bbx_delay_read_samples(self.h, _p(dst), dstformat, delay, channel, nchannels & 0xFFFFFFFF, nframes)

    def ReadSample(self, channel, delay):
        return lib().bbx_delay_read_sample(self.h, channel, delay)

    def raw(self):
        n = self.GetChannels() * self.GetLength() * FMT_BYTES[self.GetFormat()]
        out = np.zeros(n, dtype=np.uint8)
        got = lib().bbx_delay_copy_buffer(self.h, _p(out), n)
        if got != n:
            raise BbxError("bbx_delay_copy_buffer returned %d of %d bytes" % (got, n))
        return out


class SoundRingBuffer(SoundDelayBuffer):
    """SoundRingBuffer (src/SoundDelayBuffer.h:105-181): the delay buffer with a read position limiting reads and writes."""

    def __init__(self):
        h = vp()
        _check(lib().bbx_ring_create(C.byref(h)))
        self.h = h

    def GetReadPosition(self):
        return lib().bbx_ring_get_read_position(self.h)

    def GetReadFramesAvailable(self):
        return lib().bbx_ring_get_read_frames_available(self.h)

    def GetWriteFramesAvailable(self):
        return lib().bbx_ring_get_write_frames_available(self.h)

    def IncrementReadPosition(self, nframes=1):
        _check(lib().bbx_ring_increment_read_position(self.h, nframes))


# ---- next row: MultilayerBuffer (src/MultilayerBuffer.h) -------------------------------------
class MultilayerBuffer:
    def __init__(self, channels, layers):
        h = vp()
        _check(lib().bbx_mlb_create(channels, layers, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_mlb_destroy(self.h)
            self.h = None

    __del__ = close

    def GetChannels(self):
        return lib().bbx_mlb_get_channels(self.h)

    def GetLayers(self):
        return lib().bbx_mlb_get_layers(self.h)

    def GetAvailableFrames(self):
        return lib().bbx_mlb_get_available_frames(self.h)

    def WriteLayer(self, layer, src, srcchannel, nsrcchannels, ndstchannel, nchannels, nframes):
        _check(lib().bbx_mlb_write_layer(self.h, layer, _p(src), srcchannel, nsrcchannels, ndstchannel,
                                         nchannels & 0xFFFFFFFF, nframes))

    def ReadBuffer(self, srcchannel, dst, dstchannel, ndstchannels, nchannels, nframes, overwrite=True):
        return lib().bbx_mlb_read_buffer(self.h, srcchannel, _p(dst), dstchannel, ndstchannels, nchannels & 0xFFFFFFFF,
                                         nframes, int(overwrite))


# ---- next row: BiQuadCoeffs / BiQuad (src/BiQuad.h) -----------------------------------------
BIQUAD_FLAT, BIQUAD_LPF6, BIQUAD_HPF6, BIQUAD_LPF12, BIQUAD_HPF12, BIQUAD_BPF, BIQUAD_NOTCH, BIQUAD_PEQ, BIQUAD_LSH, \
    BIQUAD_HSH = range(10)


def BiQuadCalcCoeffs(ftype, freq, fs, gain=0.0, bandwidth=1.0):
    """BiQuadCoeffs(type, freq, fs, gain, bandwidth).current as [num0, num1, num2, den1, den2]."""
    out = (C.c_double * 5)()
    _check(lib().bbx_biquad_calc_coeffs(ftype, freq, fs, gain, bandwidth, out))
    return np.array(out[:], dtype=np.float64)


class BiQuadBank:
    """One BiQuadCoeffs shared by `channels` BiQuad filters, processed like BiQuad::Process(filters, ...)."""

    def __init__(self, channels):
        h = vp()
        _check(lib().bbx_biquad_create(channels, C.byref(h)))
        self.h, self.channels = h, channels

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_biquad_destroy(self.h)
            self.h = None

    __del__ = close

    def SetCoeffs(self, c5, interp_samples=0.0):
        arr = (C.c_double * 5)(*[float(v) for v in c5])
        _check(lib().bbx_biquad_set_coeffs(self.h, arr, interp_samples))

    def CalcCoeffs(self, ftype, freq, fs, gain=0.0, bandwidth=1.0, interp_time=0.0):
        _check(lib().bbx_biquad_calc(self.h, ftype, freq, fs, gain, bandwidth, interp_time))

    def Process(self, src, dst, nchannels, nsrcchannels, ndstchannels, nframes):
        _check(lib().bbx_biquad_process(self.h, _p(src), _p(dst), nchannels, nsrcchannels, ndstchannels, nframes))

    def GetState(self):
        w = np.zeros(2 * max(1, self.channels), dtype=np.float64)
        cur = np.zeros(5, dtype=np.float64)
        md = np.zeros(2, dtype=np.float64)
        _check(lib().bbx_biquad_get_state(self.h, _p(w), _p(cur), _p(md)))
        return w[:2 * self.channels], cur, md

    def Reset(self):
        _check(lib().bbx_biquad_reset(self.h))


class BiQuadFilterBank:
    """BiQuadFilterBank (src/BiQuad.h:247-353): `filters` biquads in series on each of `channels` channels, one coefficient
    object per filter; Process runs all filters in one pass over the block (k_fbank)."""

    def __init__(self, channels=0, filters=0):
        h = vp()
        _check(lib().bbx_fbank_create(channels, filters, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_fbank_destroy(self.h)
            self.h = None

    __del__ = close

    def _size(self):
        a, b = u32(0), u32(0)
        _check(lib().bbx_fbank_get_size(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def GetChannels(self):
        return self._size()[0]

    def GetFilters(self):
        return self._size()[1]

    def SetFilters(self, n):
        _check(lib().bbx_fbank_set_filters(self.h, n))

    def AddFilter(self, c5):
        _check(lib().bbx_fbank_add_filter(self.h, (C.c_double * 5)(*[float(v) for v in c5])))

    def SetChannels(self, n):
        _check(lib().bbx_fbank_set_channels(self.h, n))

    def SetCoeffs(self, filter, c5, interp_samples=0.0):
        _check(lib().bbx_fbank_set_coeffs(self.h, filter, (C.c_double * 5)(*[float(v) for v in c5]), interp_samples))

    def CalcCoeffs(self, filter, ftype, freq, fs, gain=0.0, bandwidth=1.0, interp_time=0.0):
        _check(lib().bbx_fbank_calc(self.h, filter, ftype, freq, fs, gain, bandwidth, interp_time))

    def Process(self, src, dst, nchannels, nsrcchannels, ndstchannels, nframes):
        _check(lib().bbx_fbank_process(self.h, _p(src), _p(dst), nchannels, nsrcchannels, ndstchannels, nframes))

    def ProcessDev(self, src_ptr, dst_ptr, nchannels, nsrcchannels, ndstchannels, nframes, stream=None):
        _check(lib().bbx_fbank_process_dev(self.h, src_ptr, dst_ptr, nchannels, nsrcchannels, ndstchannels, nframes, stream))

    def GetState(self, filter):
        nch = self.GetChannels()
        w = np.zeros(2 * max(1, nch), dtype=np.float64)
        cur = np.zeros(5, dtype=np.float64)
        md = np.zeros(2, dtype=np.float64)
        _check(lib().bbx_fbank_get_state(self.h, filter, _p(w), _p(cur), _p(md)))
        return w[:2 * nch], cur, md

    def Reset(self):
        _check(lib().bbx_fbank_reset(self.h))

    def Launches(self):
        n = u64(0)
        _check(lib().bbx_fbank_launches(self.h, C.byref(n)))
        return n.value


class SOFA:
    """SOFA (AES69) impulse-response set (README:77-78 lists src/SOFA.{h,cpp}; absent from the tree, so the method names
    follow the SOFA conventions, not a BBC header).  Reads the netCDF classic container; see include/bbx.h."""

    SOURCE, LISTENER, RECEIVER, EMITTER = 0, 1, 2, 3

    def __init__(self, path=None, data=None):
        self.h = None
        h = vp()
        if data is not None:
            self._data = bytes(data)
            _check(lib().bbx_sofa_open_memory(self._data, len(self._data), C.byref(h)))
        else:
            _check(lib().bbx_sofa_open(os.fsencode(path), C.byref(h)))
        self.h = h
        m, r, e, n = u32(), u32(), u32(), u32()
        _check(lib().bbx_sofa_get_sizes(h, C.byref(m), C.byref(r), C.byref(e), C.byref(n)))
        self.num_measurements, self.num_receivers, self.num_emitters, self.ir_length = m.value, r.value, e.value, n.value

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_sofa_close(self.h)
            self.h = None

    __del__ = close

    def get_samplerate(self, measurement=0):
        v = C.c_double()
        _check(lib().bbx_sofa_get_samplerate(self.h, measurement, C.byref(v)))
        return v.value

    def get_ir(self, measurement, receiver, emitter=0, n=None):
        out = np.zeros(self.ir_length if n is None else n, dtype=np.float32)
        _check(lib().bbx_sofa_get_ir(self.h, measurement, receiver, emitter, _p(out), out.size))
        return out

    def get_delay(self, measurement, receiver, emitter=0):
        v = C.c_double()
        _check(lib().bbx_sofa_get_delay(self.h, measurement, receiver, emitter, C.byref(v)))
        return v.value

    def get_position(self, which, index):
        xyz, sph = (C.c_double * 3)(), C.c_int()
        _check(lib().bbx_sofa_get_position(self.h, which, index, xyz, C.byref(sph)))
        return np.array(list(xyz)), bool(sph.value)

    def nearest_measurement(self, pos, spherical=True):
        m = u32()
        _check(lib().bbx_sofa_nearest_measurement(self.h, (C.c_double * 3)(*[float(v) for v in pos]), int(spherical), C.byref(m)))
        return m.value

    def get_attribute(self, name):
        buf = C.create_string_buffer(4096)
        _check(lib().bbx_sofa_get_attribute(self.h, name.encode(), buf, len(buf)))
        return buf.value.decode("utf-8", "replace")

    def create_filters(self, convolver, receiver, emitter=0):
        """One filter object per measurement on `convolver` (a Convolver): the bank SelectFilter chooses from."""
        arr = (vp * self.num_measurements)()
        _check(lib().bbx_sofa_create_filters(self.h, convolver.h, receiver, emitter, arr, self.num_measurements))
        out = []
        for m in range(self.num_measurements):
            f = Filter.__new__(Filter)
            f.h, f.engine = vp(arr[m]), convolver
            f.partitions = lib().bbx_filter_partitions(f.h)
            convolver._filters.append(f)
            out.append(f)
        return out


class AllPassChain:
    """AllPassFilterChain<float> (src/AllPassFilter.h): Schroeder all-pass sections over interleaved channels."""

    def __init__(self, channels, delays, coeffs):
        n = len(delays)
        d = (u32 * max(1, n))(*[int(v) for v in delays])
        c = (C.c_float * max(1, n))(*[float(v) for v in coeffs])
        h = vp()
        _check(lib().bbx_allpass_create(channels, n, d, c, C.byref(h)))
        self.h, self.channels, self.delays = h, channels, list(delays)

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_allpass_destroy(self.h)
            self.h = None

    __del__ = close

    def Process(self, src, dst, srcchannel, nsrcchannels, dstchannel, ndstchannels, nframes):
        _check(lib().bbx_allpass_process(self.h, _p(src), _p(dst), srcchannel, nsrcchannels, dstchannel, ndstchannels, nframes))

    def GetState(self, f):
        ring = np.zeros(self.channels * self.delays[f], dtype=np.float32)
        pos = lib().bbx_allpass_get_state(self.h, f, _p(ring), ring.size)
        return ring, pos


class BiQuadCascadeBank:
    """BiQuadCascade (src/BiQuad.h:373-792), one independent cascade of up to 12 float biquads per channel."""

    def __init__(self, channels, numfilters, vectorise=True, unroll=True):
        h = vp()
        _check(lib().bbx_cascade_create(channels, numfilters, int(vectorise), int(unroll), C.byref(h)))
        self.h, self.channels, self.numfilters = h, channels, numfilters

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_cascade_destroy(self.h)
            self.h = None

    __del__ = close

    def SetCoefficients(self, coeffs, channel=None):
        """(g, b1[0], b2[0], a1[0], a2[0], b1[1], ...): 4 * numfilters + 1 floats; channel None = every cascade."""
        c = np.ascontiguousarray(coeffs, dtype=np.float32)
        _check(lib().bbx_cascade_set_coefficients(self.h, 0xFFFFFFFF if channel is None else channel, _p(c), c.size))

    def Reset(self):
        _check(lib().bbx_cascade_reset(self.h))

    def ProcessCascade(self, src, dst, nframes, interleaved=True):
        _check(lib().bbx_cascade_process(self.h, _p(src), _p(dst), nframes, int(interleaved)))

    def GetState(self, channel):
        a = [np.zeros(12, dtype=np.float32) for _ in range(4)]
        last = np.zeros(1, dtype=np.float32)
        info = lib().bbx_cascade_get_state(self.h, channel, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(last))
        return a[0], a[1], a[2], a[3], last, info


# ---- a12-a14 ------------------------------------------------------------------------------
class Filter:
    """Immutable filter object: an impulse response partitioned at the engine's block size."""

    def __init__(self, engine, ir):
        ir = np.ascontiguousarray(ir, dtype=np.float32)
        h = vp()
        _check(lib().bbx_filter_create(engine.h, _p(ir), ir.size, C.byref(h)))
        self.h = h
        self.engine = engine
        self.partitions = lib().bbx_filter_partitions(h)

    def Spectra(self):
        """complex64 [partitions][B]: R2C_2B of each zero-padded partition / 2B, bin 0 = (DC, Nyquist)."""
        out = np.zeros((self.partitions, self.engine.block_size, 2), dtype=np.float32)
        _check(lib().bbx_filter_read_spectra(self.h, _p(out), out.size))
        return out.view(np.complex64)[..., 0]

    def close(self):
        """Release the spectra.  A filter that a path still has selected is refused by the library (BbxError); it
        is released together with its engine at the latest."""
        if getattr(self, "h", None) and getattr(self.engine, "h", None):
            _check(lib().bbx_filter_destroy(self.h))
            if self in self.engine._filters:
                self.engine._filters.remove(self)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Convolver:
    """Multichannel partitioned convolver on one GPU (Convolver, README:43-44)."""

    def __init__(self, block_size, max_partitions, n_inputs, n_outputs=0, n_paths=0, mode=MODE_PER_CHANNEL,
                 max_blocks=1, max_delay=0, fractional_delay=False, ring_length=0, device=0, mac_ctas_per_sm=0,
                 mac_l2_keep_16ths=0, mac_time_tile=0, mimo_tensor=0, mimo_shard_world=0, mimo_shard_rank=0):
        cfg = Config()
        cfg.device = device
        cfg.block_size = block_size
        cfg.max_partitions = max_partitions
        cfg.n_inputs = n_inputs
        cfg.n_outputs = n_outputs
        cfg.n_paths = n_paths
        cfg.mode = mode
        cfg.max_blocks = max_blocks
        cfg.max_delay = max_delay
        cfg.fractional_delay = int(fractional_delay)
        cfg.ring_length = ring_length
        cfg.mac_ctas_per_sm = mac_ctas_per_sm
        cfg.mac_l2_keep_16ths = mac_l2_keep_16ths
        cfg.mac_time_tile = mac_time_tile
        cfg.mimo_tensor = mimo_tensor
        cfg.mimo_shard_world = mimo_shard_world
        cfg.mimo_shard_rank = mimo_shard_rank
        h = vp()
        _check(lib().bbx_engine_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.block_size = block_size
        self.mode = mode
        self.n_inputs = n_inputs
        self.n_outputs = n_inputs if mode == MODE_PER_CHANNEL else n_outputs
        if mimo_shard_world > 1:
            self.n_outputs = n_outputs // mimo_shard_world  # PCM channels written by this rank
        self.n_paths = n_inputs if mode == MODE_PER_CHANNEL else (n_inputs * n_outputs if mode == MODE_MIMO else n_paths)
        self.max_blocks = max(1, max_blocks)
        self.ring_length = lib().bbx_engine_get_ring_length(h)
        self._filters = []

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_engine_destroy(self.h)  # releases every filter the engine still owns
            for f in self._filters:
                f.h = None
            self._filters = []
            self.h = None

    __del__ = close

    def CreateFilter(self, ir):
        f = Filter(self, ir)
        self._filters.append(f)
        return f

    def SetRoute(self, path, inp, out, gain=1.0):
        _check(lib().bbx_set_route(self.h, path, inp, out, gain))

    def SelectFilter(self, path, flt, delay=0.0, crossfade=False):
        _check(lib().bbx_set_filter(self.h, path, flt.h if flt is not None else None, int(crossfade), delay))

    def SelectFilters(self, paths, filters, delays=None, crossfade=None):
        """SelectFilter for many paths in one call (all validated before any is latched)."""
        n = len(paths)
        pa = (u32 * n)(*[int(p) for p in paths])
        fa = (vp * n)(*[(f.h if f is not None else None) for f in filters])
        xa = (C.c_int * n)(*[int(bool(x)) for x in crossfade]) if crossfade is not None else None
        da = (C.c_double * n)(*[float(d) for d in delays]) if delays is not None else None
        _check(lib().bbx_set_filters(self.h, n, pa, fa, xa, da))

    def Convolve(self, inp, infmt, in_channels, outfmt, out_channels, nframes, in_be=False, out_be=False, out=None):
        """Host buffers.  inp: interleaved PCM bytes/typed array; returns the output byte buffer."""
        inp = np.ascontiguousarray(inp).view(np.uint8).reshape(-1)
        assert inp.size == nframes * in_channels * FMT_BYTES[infmt], "input size does not match the geometry"
        if out is None:
            out = np.zeros(nframes * out_channels * FMT_BYTES[outfmt], dtype=np.uint8)
        _check(lib().bbx_process(self.h, _p(inp), infmt, int(in_be), in_channels, _p(out), outfmt, int(out_be),
                                 out_channels, nframes))
        return out

    def ConvolveHostPtr(self, in_ptr, infmt, in_channels, out_ptr, outfmt, out_channels, nframes):
        """Raw host pointers (e.g. pinned buffers); synchronous, copies included."""
        _check(lib().bbx_process(self.h, vp(in_ptr), infmt, 0, in_channels, vp(out_ptr), outfmt, 0, out_channels, nframes))

    def BlockLatency(self, in_ptr, infmt, in_channels, out_ptr, outfmt, out_channels, nframes, ncalls, warmup=100):
        """Host time (us) of each of ncalls synchronous bbx_process calls on the same buffers, measured inside libbbx
        (bbx_block_latency): what a C / C++ caller sees, without the ctypes call around every block."""
        us = np.empty(ncalls, dtype=np.float64)
        _check(lib().bbx_block_latency(self.h, vp(in_ptr), infmt, 0, in_channels, vp(out_ptr), outfmt, 0, out_channels, nframes,
                                       ncalls, warmup, us.ctypes.data_as(C.POINTER(C.c_double))))
        return us

    def ConvolveHostPtrAsync(self, in_ptr, infmt, in_channels, out_ptr, outfmt, out_channels, nframes):
        """Raw PINNED host pointers; returns after enqueueing (H2D, kernels, D2H on separate streams).  The buffers
        must stay untouched until Sync()."""
        _check(lib().bbx_process_async(self.h, vp(in_ptr), infmt, 0, in_channels, vp(out_ptr), outfmt, 0, out_channels,
                                       nframes))

    def ConvolveDev(self, in_ptr, infmt, in_channels, out_ptr, outfmt, out_channels, nframes, in_be=False, out_be=False):
        """Device pointers (ints); asynchronous on the engine stream."""
        _check(lib().bbx_process_dev(self.h, vp(in_ptr), infmt, int(in_be), in_channels, vp(out_ptr), outfmt, int(out_be),
                                     out_channels, nframes))

    def Sync(self):
        _check(lib().bbx_engine_sync(self.h))

    # measurement hooks
    def timer_start(self):
        _check(lib().bbx_engine_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        _check(lib().bbx_engine_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return lib().bbx_engine_launch_count(self.h)

    def mac_kernel_name(self):
        return lib().bbx_engine_mac_kernel(self.h).decode()

    def profile_mac(self, enable=True):
        _check(lib().bbx_engine_profile_mac(self.h, int(enable)))

    def mac_time(self):
        ms, n, units, nbytes = C.c_float(0), u64(0), u64(0), u64(0)
        _check(lib().bbx_engine_mac_time(self.h, C.byref(ms), C.byref(n), C.byref(units), C.byref(nbytes)))
        return {"ms": ms.value, "launches": n.value, "channel_blocks": units.value, "algorithmic_bytes": nbytes.value}

    def GetState(self):
        """The engine's audio state as a uint8 array (bbx_engine_get_state): checkpoint."""
        n = C.c_size_t(0)
        _check(lib().bbx_engine_state_size(self.h, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        _check(lib().bbx_engine_get_state(self.h, _p(buf), buf.size))
        return buf

    def SetState(self, state):
        """Resume from a GetState() blob; this engine must hold the same filters in the same order."""
        buf = np.ascontiguousarray(state, dtype=np.uint8)
        _check(lib().bbx_engine_set_state(self.h, _p(buf), buf.size))

    def io_trace(self, calls):
        """Trace the next `calls` ConvolveHostPtrAsync calls (bbx_engine_io_trace)."""
        _check(lib().bbx_engine_io_trace(self.h, calls))

    def io_trace_read(self, cap=4096):
        """[n][6] ms: h2d start/end, kernels start/end, d2h start/end of every traced call."""
        buf = (C.c_float * (6 * cap))()
        n = u32(0)
        _check(lib().bbx_engine_io_trace_read(self.h, buf, cap, C.byref(n)))
        return [[buf[6 * i + j] for j in range(6)] for i in range(n.value)]

    def exchange_time(self):
        """input-sharded MIMO, while profile_mac is on: device time of the exchange step, exchanges, bytes sent to peers"""
        ms, n, nbytes = C.c_float(0), u64(0), u64(0)
        _check(lib().bbx_engine_exchange_time(self.h, C.byref(ms), C.byref(n), C.byref(nbytes)))
        return {"ms": ms.value, "exchanges": n.value, "bytes_sent": nbytes.value}

    def set_tuning(self, ctas_per_sm=0, l2_keep_16ths=0, time_tile=0):
        """0 = leave as is; time_tile=1 forces the streaming MAC, l2_keep_16ths > 16 switches the hints off."""
        _check(lib().bbx_engine_set_tuning(self.h, ctas_per_sm, l2_keep_16ths, time_tile))

    def PeerExport(self):
        """64-byte CUDA IPC handle of this rank's receive buffer (input-sharded MIMO engine, peer-memory mixdown)."""
        h = (C.c_uint8 * 64)()
        _check(lib().bbx_engine_peer_export(self.h, h))
        return bytes(h)

    def PeerAttach(self, handles):
        """handles: the PeerExport() results of all ranks, in rank order."""
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(lib().bbx_engine_peer_attach(self.h, buf))

    def set_mixdown_kernel(self, per_output):
        """True keeps many-path mixdowns on the per-output kernel instead of k_pcm_out_mix (identical bytes)."""
        _check(lib().bbx_engine_set_mixdown_kernel(self.h, int(bool(per_output))))

    def set_direct_io(self, max_bytes):
        """Largest PCM buffer (bytes) for which host calls with pinned buffers bypass the copy engines; 0 disables."""
        _check(lib().bbx_engine_set_direct_io(self.h, int(max_bytes)))

    def direct_calls(self):
        return int(lib().bbx_engine_direct_calls(self.h))

    def set_fused(self, enable, max_partitions=0):
        """single-launch latency path (k_block_fused) for one-block calls of PER_CHANNEL / ROUTED engines; False keeps the
        multi-kernel path (identical bytes)"""
        _check(lib().bbx_engine_set_fused(self.h, int(bool(enable)), int(max_partitions)))

    def fused_calls(self):
        return int(lib().bbx_engine_fused_calls(self.h))

    def SetComm(self, comm):
        """Attach the communicator of the input-sharded MIMO engine (None detaches)."""
        _check(lib().bbx_engine_set_comm(self.h, comm.h if comm is not None else None))
        self._comm = comm

    def tensor_status(self):
        """(launches of the tensor-core MIMO kernel so far, device status word: 0 = ok)."""
        n, st = u64(0), C.c_int(0)
        _check(lib().bbx_engine_tensor_status(self.h, C.byref(n), C.byref(st)))
        return n.value, st.value

    def tensor_trace(self, n_ctas, enable=None):
        """enable=True/False switches the per-CTA role trace of k_mimo_tc; otherwise returns uint64[n_ctas][16]."""
        if enable is not None:
            _check(lib().bbx_engine_tensor_trace(self.h, None, n_ctas if enable else 0))
            return None
        out = np.zeros((n_ctas, 16), dtype=np.uint64)
        _check(lib().bbx_engine_tensor_trace(self.h, _p(out), n_ctas))
        return out

    def flush_l2(self, nbytes=256 << 20):
        _check(lib().bbx_engine_flush_l2(self.h, nbytes))


def probe_fp32_tflops(device=0, seconds=0.5):
    """(burst, sustained) TFLOP/s of a pure packed-FMA kernel on this GPU (FP32 roofline denominator)."""
    b, s_ = C.c_float(0), C.c_float(0)
    _check(lib().bbx_probe_fp32_tflops(device, seconds, C.byref(b), C.byref(s_)))
    return b.value, s_.value


def probe_launch_sync(device=0, ncalls=2000, warmup=200):
    """Host time (us) of ncalls round trips "one trivial kernel + cudaStreamSynchronize", timed inside libbbx."""
    us = np.empty(ncalls, dtype=np.float64)
    _check(lib().bbx_probe_launch_sync(device, ncalls, warmup, us.ctypes.data_as(C.POINTER(C.c_double))))
    return us


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the other ranks receive it out of band)."""
    buf = (u8 * 128)()
    _check(lib().bbx_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """Communicator of the input-sharded MIMO engine (NCCL, loaded at run time by libbbx)."""

    def __init__(self, world, rank, unique_id, device=0):
        assert len(unique_id) == 128
        buf = (u8 * 128).from_buffer_copy(unique_id)
        h = vp()
        _check(lib().bbx_comm_create(world, rank, buf, device, C.byref(h)))
        self.h, self.world, self.rank = h, world, rank

    def close(self):
        if getattr(self, "h", None):
            lib().bbx_comm_destroy(self.h)
            self.h = None


class BlockConvolver:
    """Single-channel partitioned convolution (BlockConvolver, README:38-39): Convolve() one block."""

    def __init__(self, block_size, max_partitions, device=0):
        self.engine = Convolver(block_size, max_partitions, 1, device=device)
        self.block_size = block_size

    def CreateFilter(self, ir):
        return self.engine.CreateFilter(ir)

    def SetFilter(self, flt, crossfade=False):
        self.engine.SelectFilter(0, flt, 0.0, crossfade)

    def Convolve(self, block):
        block = np.ascontiguousarray(block, dtype=np.float32)
        assert block.size == self.block_size
        out = np.zeros(self.block_size, dtype=np.float32)
        _check(lib().bbx_blockconvolver_convolve(self.engine.h, _p(block), _p(out)))
        return out

    def close(self):
        self.engine.close()


class PinnedBuffer:
    """Pinned host memory from bbx_host_alloc, viewed as a numpy uint8 array."""

    def __init__(self, nbytes):
        p = vp()
        _check(lib().bbx_host_alloc(C.byref(p), nbytes))
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            lib().bbx_host_free(vp(self.ptr))
            self.ptr = None

    __del__ = close
