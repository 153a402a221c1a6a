"""The single-launch latency path (k_block_fused, bbx_engine_set_fused): streaming calls (one block) of PER_CHANNEL and
ROUTED engines with short filters.  Contract: the BYTES of the multi-kernel path (k_pcm_in -> k_rfft -> k_fdl_mac ->
k_irfft -> k_pcm_out), for every block size, sample format, delay mode, crossfaded / hard switch and silent path, and the
same engine state afterwards (calls of other sizes continue seamlessly); plus the oracle as an independent check."""
import numpy as np
import pytest

import cpulibs as cl
from convkit import GpuDriver, OracleDriver, interleave, make_ir, make_noise
from parity import assert_float_parity

pytestmark = pytest.mark.gpu


def run_script(bbx, fused, B, nch, P, calls, infmt, outfmt, in_be=False, out_be=False, fractional=False, max_delay=40, seed=0,
               mode=cl.MODE_PER_CHANNEL, n_out=0, n_paths=0):
    """calls: list of block counts; filters / delays switch before some calls (hard and crossfaded), one path goes silent"""
    rng = np.random.default_rng(seed)
    tmax = max(calls)
    g = GpuDriver(bbx, B, P, nch, n_outputs=n_out, n_paths=n_paths, mode=mode, max_blocks=tmax, max_delay=max_delay,
                  fractional_delay=fractional)
    g.eng.set_fused(fused)
    npaths = n_paths if mode == cl.MODE_ROUTED else nch
    nouts = n_out if mode == cl.MODE_ROUTED else nch
    lens = [int(rng.integers(B // 2, P * B + 1)) for _ in range(npaths + 3)]
    fl = [g.filter(make_ir(100 + seed * 50 + k, lens[k])) for k in range(npaths + 3)]
    if mode == cl.MODE_ROUTED:
        for p in range(npaths):
            g.route(p, p % nch, (p * 7) % nouts, [1.0, 0.5, 0.0, -0.25][p % 4])
    for p in range(npaths):
        d = float(rng.integers(0, max_delay)) + (0.37 * p if fractional else 0.0)
        g.select(p, fl[p], delay=min(d, float(max_delay)))
    bps_in, bps_out = cl.FMT_BYTES[infmt], cl.FMT_BYTES[outfmt]
    outs = []
    for ci, nb in enumerate(calls):
        if ci == 2:  # crossfaded filter + delay switch on path 0, hard switch on path 1, path 2 silenced
            g.select(0, fl[npaths], delay=11.5 if fractional else 11.0, crossfade=True)
            if npaths > 1:
                g.select(1, fl[npaths + 1], delay=3.0)
            if npaths > 2:
                g.select(2, None)
        if ci == 4:  # delay-only crossfade
            g.select(0, fl[npaths], delay=2.0, crossfade=True)
        n = nb * B
        if infmt >= cl.FMT_FLOAT:
            x = rng.uniform(-0.5, 0.5, n * nch)
            raw = (x.astype(np.float32) if infmt == cl.FMT_FLOAT else x).view(np.uint8).copy()
            if in_be:
                raw = raw.reshape(-1, bps_in)[:, ::-1].reshape(-1).copy()
        else:
            raw = rng.integers(0, 256, n * nch * bps_in, dtype=np.uint8)
        outs.append(g.process(raw, infmt, nch, outfmt, nouts, n, in_be, out_be).copy())
        assert outs[-1].size == n * nouts * bps_out
    fused_calls = g.eng.fused_calls()
    g.close()
    return np.concatenate(outs), fused_calls


@pytest.mark.parametrize("B", [64, 128, 256, 512, 1024, 2048, 4096])
def test_fused_bytes_equal_multikernel_every_block_size(bbx, B):
    calls = [1, 1, 1, 3, 1, 1, 2, 1]  # single blocks (fused) between longer calls (multi-kernel): the state carries over
    nch = 5 if B <= 1024 else 2
    a, na = run_script(bbx, True, B, nch, 4, calls, cl.FMT_FLOAT, cl.FMT_FLOAT, seed=B)
    b, nb = run_script(bbx, False, B, nch, 4, calls, cl.FMT_FLOAT, cl.FMT_FLOAT, seed=B)
    assert na == calls.count(1) and nb == 0
    assert np.array_equal(a, b)


@pytest.mark.parametrize("infmt,outfmt,in_be,out_be,fractional", [
    (cl.FMT_24, cl.FMT_24, False, False, True),    # C4: int24 in / out, fractional delays
    (cl.FMT_16, cl.FMT_32, True, False, False),
    (cl.FMT_32, cl.FMT_16, False, True, True),
    (cl.FMT_DOUBLE, cl.FMT_FLOAT, False, False, False),
    (cl.FMT_FLOAT, cl.FMT_DOUBLE, True, True, True),
])
def test_fused_bytes_equal_multikernel_formats_and_delays(bbx, infmt, outfmt, in_be, out_be, fractional):
    calls = [1, 1, 1, 1, 1, 2, 1, 1]
    kw = dict(in_be=in_be, out_be=out_be, fractional=fractional, seed=7)
    a, na = run_script(bbx, True, 512, 33, 8, calls, infmt, outfmt, **kw)   # 33 channels: a ragged last CTA
    b, _ = run_script(bbx, False, 512, 33, 8, calls, infmt, outfmt, **kw)
    assert na == calls.count(1)
    assert np.array_equal(a, b)


def test_fused_routed_mixdown_bytes_equal_multikernel(bbx):
    """C2's shape: sources fan out to two ears each, zero gains, ITD delays; the mixdown stays a second launch"""
    calls = [1, 1, 1, 1, 1, 4, 1]
    kw = dict(mode=cl.MODE_ROUTED, n_out=2, n_paths=24, max_delay=48, seed=3)
    a, na = run_script(bbx, True, 256, 12, 2, calls, cl.FMT_FLOAT, cl.FMT_FLOAT, **kw)
    b, _ = run_script(bbx, False, 256, 12, 2, calls, cl.FMT_FLOAT, cl.FMT_FLOAT, **kw)
    assert na == calls.count(1)
    assert np.array_equal(a, b)


def test_fused_long_filters_stay_on_the_multikernel_path(bbx):
    a, na = run_script(bbx, True, 128, 3, 40, [1, 1, 2, 1], cl.FMT_FLOAT, cl.FMT_FLOAT, seed=5)
    assert na <= 4  # paths longer than 32 partitions are not fused (random lengths: some calls may qualify after a switch)


def test_fused_vs_oracle(bbx, orc):
    """the fused path against the CPU oracle block by block (C1's shape: 8192 taps, B = 1024)"""
    B, L, nch, nblk = 1024, 8192, 2, 12
    g = GpuDriver(bbx, B, 8, nch, max_blocks=1)
    o = OracleDriver(B, 8, nch, max_blocks=1)
    irs = [make_ir(2000 + c, L) for c in range(nch)]
    for c in range(nch):
        g.select(c, g.filter(irs[c]))
        o.select(c, o.filter(irs[c]))
    xs = interleave([make_noise(1000 + c, nblk * B) for c in range(nch)])
    yg = np.concatenate([g.process(xs[i * B:(i + 1) * B], cl.FMT_FLOAT, nch, cl.FMT_FLOAT, nch, B).view(np.float32).reshape(B, nch)
                         for i in range(nblk)])
    yo = np.concatenate([o.process(xs[i * B:(i + 1) * B], cl.FMT_FLOAT, nch, cl.FMT_FLOAT, nch, B).view(np.float32).reshape(B, nch)
                         for i in range(nblk)])
    assert g.eng.fused_calls() == nblk
    g.close()
    for c in range(nch):
        assert_float_parity(yg[:, c], yo[:, c], "fused vs oracle ch %d" % c)
