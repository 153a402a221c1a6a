"""Convolver test kit (TEST INFRASTRUCTURE): synthetic inputs per SURVEY.md 8(d) and one driver surface
over the CPU oracle convolver and the CUDA product, so the same scenario runs on both."""
import numpy as np

import cpulibs as cl


def make_ir(seed, L):
    """rng(seed).standard_normal(L) * exp(-6.9 n / L), L2-normalised, fp32."""
    rng = np.random.default_rng(seed)
    h = rng.standard_normal(L) * np.exp(-6.9 * np.arange(L) / L)
    h /= np.sqrt((h ** 2).sum())
    return h.astype(np.float32)


def make_noise(seed, n):
    """uniform white noise in [-1, 1), fp32."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n).astype(np.float32)


def interleave(chans):
    """list of per-channel arrays -> [frames][channels] C-contiguous."""
    return np.ascontiguousarray(np.stack(chans, axis=1))


class OracleDriver:
    name = "oracle"

    def __init__(self, block, max_partitions, n_inputs, n_outputs=0, n_paths=0, mode=cl.MODE_PER_CHANNEL, max_blocks=1,
                 max_delay=0, fractional_delay=False, ring_length=0):
        self.orc = cl.oracle()
        if ring_length == 0:  # same default as libbbx (bbx_engine_create)
            need = max_delay + 14 + (max(1, max_blocks) + 1) * block
            ring_length = -(-need // block) * block
        self.ring_length = ring_length
        self.block = block
        self.cv = self.orc.convolver(block=block, max_partitions=max_partitions, n_inputs=n_inputs,
                                     n_outputs=n_outputs or n_inputs, n_paths=n_paths, mode=mode, ring_len=ring_length,
                                     fractional_delay=fractional_delay)

    def filter(self, ir):
        return self.orc.filter(ir, self.block)

    def route(self, path, inp, out, gain=1.0):
        self.cv.set_route(path, inp, out, gain)

    def select(self, path, f, delay=0.0, crossfade=False):
        self.cv.set_filter(path, f, crossfade, delay)

    def process(self, x, infmt, in_channels, outfmt, out_channels, nframes, in_be=False, out_be=False):
        return self.cv.process(x, infmt, in_channels, outfmt, out_channels, nframes, in_be, out_be)

    def close(self):
        pass


class GpuDriver:
    name = "gpu"

    def __init__(self, bbx, block, max_partitions, n_inputs, n_outputs=0, n_paths=0, mode=cl.MODE_PER_CHANNEL,
                 max_blocks=1, max_delay=0, fractional_delay=False, ring_length=0, **kw):
        self.eng = bbx.Convolver(block, max_partitions, n_inputs, n_outputs, n_paths, mode, max_blocks, max_delay,
                                 fractional_delay, ring_length, **kw)
        self.ring_length = self.eng.ring_length
        self.block = block

    def filter(self, ir):
        return self.eng.CreateFilter(ir)

    def route(self, path, inp, out, gain=1.0):
        self.eng.SetRoute(path, inp, out, gain)

    def select(self, path, f, delay=0.0, crossfade=False):
        self.eng.SelectFilter(path, f, delay, crossfade)

    def process(self, x, infmt, in_channels, outfmt, out_channels, nframes, in_be=False, out_be=False):
        return self.eng.Convolve(x, infmt, in_channels, outfmt, out_channels, nframes, in_be, out_be)

    def close(self):
        self.eng.close()


def run_float(driver, x, nframes_per_call):
    """x: float32 [frames][channels]; returns float32 [frames][n_out] processing in calls of the given size."""
    frames, ch = x.shape
    outs = []
    pos = 0
    sizes = nframes_per_call if isinstance(nframes_per_call, (list, tuple)) else None
    i = 0
    while pos < frames:
        n = sizes[i % len(sizes)] if sizes else nframes_per_call
        n = min(n, frames - pos)
        o = driver.process(x[pos:pos + n], cl.FMT_FLOAT, ch, cl.FMT_FLOAT, driver_out_channels(driver), n)
        outs.append(o.view(np.float32).reshape(n, -1))
        pos += n
        i += 1
    return np.concatenate(outs, axis=0)


def driver_out_channels(driver):
    if driver.name == "oracle":
        return driver.cv.n_outputs
    return driver.eng.n_outputs
