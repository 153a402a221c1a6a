"""Adapter giving the CUDA product (through the C ABI) the same call surface as cpulibs.CpuLib, so one
test body checks the oracle on CPU and libbbx on the GPU."""
import numpy as np


class GpuLib:
    def __init__(self, bbx):
        self.b = bbx

    def bits_per_sample(self, fmt):
        return self.b.GetBitsPerSample(fmt)

    def bytes_per_sample(self, fmt):
        return self.b.GetBytesPerSample(fmt)

    def sanity(self, *a, allow=True):
        return self.b.BlockTransferSanityChecks(*a, allowsinglechannel=allow)

    def transfer(self, src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel,
                 dst_channels, nchannels, nframes):
        self.b.TransferSamples(src, srctype, src_be, src_channel, src_channels, dst, dsttype, dst_be, dst_channel,
                               dst_channels, nchannels, nframes)

    def transfer_linear(self, src, srctype, dst, dsttype, nsamples):
        self.b.TransferSamplesLinear(src, srctype, dst, dsttype, nsamples)

    def mix(self, src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul=1.0):
        self.b.MixSamples(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, mul)

    def mix_interp(self, src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes, state, inc):
        self.b.MixSamples(src, src_channel, src_channels, dst, dst_channel, dst_channels, nchannels, nframes,
                          interp=state, inc=inc)

    def interp_step(self, state, inc, nsteps):
        self.b.InterpolatorStep(state, inc, nsteps)

    def frac_additional(self):
        return self.b.FractionalSampleAdditionalDelayRequired()

    def frac(self, buffer, channel, channels, length, pos):
        return self.b.FractionalSample(buffer, channel, channels, length, np.asarray(pos, dtype=np.float64))

    def delay(self, ring=False):
        return GpuDelay(self.b, ring)

    def multilayer(self, channels, layers):
        return GpuMultilayer(self.b, channels, layers)

    def biquad_coeffs(self, ftype, freq, fs, gain=0.0, bandwidth=1.0):
        return self.b.BiQuadCalcCoeffs(ftype, freq, fs, gain, bandwidth)

    def biquad(self, channels):
        return GpuBiquad(self.b, channels)

    def fbank(self, channels, filters):
        return GpuFbank(self.b, channels, filters)

    def allpass(self, channels, delays, coeffs):
        return GpuAllpass(self.b, channels, delays, coeffs)

    def cascade(self, channels, numfilters, vectorise=True, unroll=True):
        return GpuCascade(self.b, channels, numfilters, vectorise, unroll)


class GpuCascade:
    def __init__(self, b, channels, numfilters, vectorise, unroll):
        self.c = b.BiQuadCascadeBank(channels, numfilters, vectorise, unroll)

    def close(self):
        self.c.close()

    def set_coefficients(self, coeffs, channel=None):
        try:
            self.c.SetCoefficients(coeffs, channel)
            return True
        except Exception:
            return False

    def reset(self):
        self.c.Reset()

    def process(self, src, dst, nframes, interleaved=True):
        self.c.ProcessCascade(src, dst, nframes, interleaved)

    def state(self, channel):
        return self.c.GetState(channel)


class GpuAllpass:
    def __init__(self, b, channels, delays, coeffs):
        self.a = b.AllPassChain(channels, delays, coeffs)

    def close(self):
        self.a.close()

    def process(self, src, dst, srcchannel, nsrc, dstchannel, ndst, nframes):
        self.a.Process(src, dst, srcchannel, nsrc, dstchannel, ndst, nframes)

    def state(self, f):
        return self.a.GetState(f)


class GpuBiquad:
    def __init__(self, b, channels):
        self.q = b.BiQuadBank(channels)

    def close(self):
        self.q.close()

    def set_coeffs(self, c5, interp_samples=0.0):
        self.q.SetCoeffs(c5, interp_samples)

    def calc(self, ftype, freq, fs, gain=0.0, bandwidth=1.0, interp_time=0.0):
        self.q.CalcCoeffs(ftype, freq, fs, gain, bandwidth, interp_time)

    def process(self, src, dst, nchannels, nsrc, ndst, nframes):
        self.q.Process(src, dst, nchannels, nsrc, ndst, nframes)

    def state(self):
        return self.q.GetState()

    def reset(self):
        self.q.Reset()


class GpuFbank:
    def __init__(self, b, channels, filters):
        self.q = b.BiQuadFilterBank(channels, filters)

    def close(self):
        self.q.close()

    def set_filters(self, n):
        self.q.SetFilters(n)

    def add_filter(self, c5):
        self.q.AddFilter(c5)

    def set_channels(self, n):
        self.q.SetChannels(n)

    def set_coeffs(self, filter, c5, interp_samples=0.0):
        self.q.SetCoeffs(filter, c5, interp_samples)

    def calc(self, filter, ftype, freq, fs, gain=0.0, bandwidth=1.0, interp_time=0.0):
        self.q.CalcCoeffs(filter, ftype, freq, fs, gain, bandwidth, interp_time)

    def process(self, src, dst, nchannels, nsrc, ndst, nframes):
        self.q.Process(src, dst, nchannels, nsrc, ndst, nframes)

    def state(self, filter):
        return self.q.GetState(filter)

    def reset(self):
        self.q.Reset()


class GpuMultilayer:
    def __init__(self, b, channels, layers):
        self.m = b.MultilayerBuffer(channels, layers)

    def close(self):
        self.m.close()

    def write_layer(self, layer, src, srcchannel, nsrcchannels, dstchannel, nchannels, nframes):
        self.m.WriteLayer(layer, src, srcchannel, nsrcchannels, dstchannel, nchannels, nframes)

    def available(self):
        return self.m.GetAvailableFrames()

    def read(self, srcchannel, dst, dstchannel, ndstchannels, nchannels, nframes, overwrite=True):
        return self.m.ReadBuffer(srcchannel, dst, dstchannel, ndstchannels, nchannels, nframes, overwrite)


class GpuDelay:
    def __init__(self, b, ring=False):
        self.d = b.SoundRingBuffer() if ring else b.SoundDelayBuffer()

    read_position = property(lambda s: s.d.GetReadPosition())
    read_available = property(lambda s: s.d.GetReadFramesAvailable())
    write_available = property(lambda s: s.d.GetWriteFramesAvailable())

    def increment_read(self, nframes):
        self.d.IncrementReadPosition(nframes)

    def close(self):
        self.d.close()

    def set_size(self, chans, length, fmt=4):
        self.d.SetSize(chans, length, fmt)

    channels = property(lambda s: s.d.GetChannels())
    length = property(lambda s: s.d.GetLength())
    write_position = property(lambda s: s.d.GetWritePosition())
    format = property(lambda s: s.d.GetFormat())

    def write(self, src, srcformat, channel, nchannels, nframes):
        return self.d.WriteSamples(src, srcformat, channel, nchannels, nframes)

    def increment(self, nframes):
        self.d.IncrementWritePosition(nframes)

    def read(self, dst, dstformat, delay, channel, nchannels, nframes):
        return self.d.ReadSamples(dst, dstformat, delay, channel, nchannels, nframes)

    def raw(self):
        return self.d.raw()
