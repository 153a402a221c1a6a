"""CPU tier: the oracle's UPOLS BlockConvolver / Convolver against float64 truth.

The reference's BlockConvolver/Convolver are absent from the mounted tree (README:38-51), so the convolver
oracle is pinned by the definition itself: float64 direct convolution (golden fixture conv_truth.npz generated
with numpy float64, plus oracle/direct.c) and metamorphic properties.  Tolerance: SNR >= 110 dB and max-abs
<= 1e-5 x peak (BASELINE.json north_star).
"""
import numpy as np
import pytest

import cpulibs as cl
from conftest import golden
from convkit import OracleDriver, interleave, make_ir, make_noise, run_float
from gen_golden import conv_case
from parity import assert_float_parity, s24_to_float


def test_fft_against_numpy(orc):
    rng = np.random.default_rng(3)
    for n in (8, 128, 512, 1024, 2048, 8192):
        x = rng.standard_normal(n).astype(np.float32)
        X = orc.rfft(x)
        Xn = np.fft.rfft(x.astype(np.float64))
        assert np.abs(X - Xn).max() <= 3e-7 * np.abs(Xn).max() * np.log2(n)
        y = orc.irfft(Xn.astype(np.complex64), n) / n
        assert np.abs(y - x).max() <= 2e-6


def test_blockconvolver_golden_truth(orc):
    g = golden("conv_truth.npz")
    for seed, L, B, nblk in g["cases"]:
        h, x = conv_case(int(seed), int(L), int(B), int(nblk))
        f = orc.filter(h[0], int(B))
        bc = orc.blockconv(int(B), f.partitions)
        bc.set_filter(f)
        y = np.concatenate([bc.convolve(x[0][i * B:(i + 1) * B]) for i in range(nblk)])
        assert_float_parity(y, g["y_%d" % seed], "L=%d B=%d" % (L, B))


def test_direct_c_matches_numpy(orc):
    rng = np.random.default_rng(4)
    x, h = rng.standard_normal(3000), rng.standard_normal(257)
    want = np.convolve(x, h)[:3000]
    assert np.abs(orc.direct(x, h) - want).max() < 1e-12
    assert np.abs(orc.direct(x, h, n0=2900, count=100) - want[2900:]).max() < 1e-12


def test_long_reverb_window_vs_direct(orc):
    """C3 shape (144000 taps, B=512, P=282) on one channel; direct float64 on a spot-checked window."""
    L, B, nblk = 144000, 512, 300
    h, x = make_ir(2000, L), make_noise(1000, nblk * B)
    f = orc.filter(h, B)
    assert f.partitions == 282
    bc = orc.blockconv(B, 282)
    bc.set_filter(f)
    y = np.concatenate([bc.convolve(x[i * B:(i + 1) * B]) for i in range(nblk)])
    n0 = nblk * B - 2048
    yd = orc.direct(x, h, n0=n0, count=2048)
    assert_float_parity(y[n0:], yd, "C3 window")


@pytest.mark.parametrize("B,L", [(64, 300), (256, 512), (512, 4096)])
def test_impulse_and_delayed_impulse(orc, B, L):
    P = -(-L // B)
    nblk = P + 8  # enough signal after the longest delay for a meaningful SNR
    x = make_noise(1, nblk * B)
    for tap in (0, B - 1, B, B + 1, L - 1):
        h = np.zeros(L, dtype=np.float32)
        h[tap] = 1.0
        f = orc.filter(h, B)
        bc = orc.blockconv(B, P)
        bc.set_filter(f)
        y = np.concatenate([bc.convolve(x[i * B:(i + 1) * B]) for i in range(nblk)])
        want = np.concatenate([np.zeros(tap, dtype=np.float32), x])[: x.size]
        assert_float_parity(y, want, "tap %d" % tap)


def test_zero_input_gives_exact_zeros(orc):
    B = 128
    f = orc.filter(make_ir(5, 700), B)
    bc = orc.blockconv(B, 6)
    bc.set_filter(f)
    for _ in range(4):
        assert not bc.convolve(np.zeros(B, dtype=np.float32)).any()


def test_linearity_and_block_size_invariance(orc):
    L = 1500
    h = make_ir(6, L)
    x1, x2 = make_noise(7, 4096), make_noise(8, 4096)

    def run(x, B):
        f = orc.filter(h, B)
        bc = orc.blockconv(B, f.partitions)
        bc.set_filter(f)
        return np.concatenate([bc.convolve(x[i * B:(i + 1) * B]) for i in range(x.size // B)])

    a, b, ab = run(x1, 256), run(x2, 256), run((x1 + 0.5 * x2).astype(np.float32), 256)
    assert_float_parity(ab, a.astype(np.float64) + 0.5 * b.astype(np.float64), "linearity")
    assert_float_parity(run(x1, 64), run(x1, 1024), "block-size invariance")


def test_crossfaded_filter_switch_vs_direct(orc):
    """A crossfaded switch equals two convolvers on the same history, blended with g_n = n/B for one block."""
    B, L, nblk, sw = 128, 700, 10, 4
    h1, h2 = make_ir(11, L), make_ir(12, L)
    x = make_noise(13, nblk * B)
    f1, f2 = orc.filter(h1, B), orc.filter(h2, B)
    bc = orc.blockconv(B, 6)
    bc.set_filter(f1)
    out = []
    for i in range(nblk):
        if i == sw:
            bc.set_filter(f2, crossfade=True)
        out.append(bc.convolve(x[i * B:(i + 1) * B]))
    y = np.concatenate(out)
    y1, y2 = orc.direct(x, h1), orc.direct(x, h2)
    want = y1.copy()
    g = np.arange(B) / B
    want[sw * B:(sw + 1) * B] = (1 - g) * y1[sw * B:(sw + 1) * B] + g * y2[sw * B:(sw + 1) * B]
    want[(sw + 1) * B:] = y2[(sw + 1) * B:]
    assert_float_parity(y, want, "crossfade")
    # hard switch: the new filter applies from the boundary on
    bc = orc.blockconv(B, 6)
    bc.set_filter(f1)
    out = []
    for i in range(nblk):
        if i == sw:
            bc.set_filter(f2, crossfade=False)
        out.append(bc.convolve(x[i * B:(i + 1) * B]))
    want = np.concatenate([y1[:sw * B], y2[sw * B:]])
    assert_float_parity(np.concatenate(out), want, "hard switch")


def test_null_filter_is_silence(orc):
    bc = orc.blockconv(64, 2)
    assert not bc.convolve(make_noise(1, 64)).any()


def test_convolver_per_channel_integer_delay(orc):
    B, L, nch, nblk = 64, 150, 3, 12
    irs = [make_ir(20 + c, L) for c in range(nch)]
    xs = [make_noise(30 + c, nblk * B) for c in range(nch)]
    drv = OracleDriver(B, 3, nch, max_blocks=4, max_delay=100)
    delays = [0, 7, 100]
    for c in range(nch):
        drv.select(c, drv.filter(irs[c]), delay=delays[c])
    y = run_float(drv, interleave(xs), [B, 3 * B, 4 * B])
    for c in range(nch):
        want = np.concatenate([np.zeros(delays[c]), orc.direct(xs[c], irs[c])])[: nblk * B]
        assert_float_parity(y[:, c], want, "ch %d" % c)


def test_convolver_fractional_delay_and_crossfade(orc):
    """Fractional delay = FractionalSample over the ring of the convolved stream; checked against the same
    14-tap read applied to the float64 truth, including a crossfaded delay+filter switch."""
    B, L, nblk, sw = 128, 300, 10, 5
    h1, h2 = make_ir(41, L), make_ir(42, L)
    x = make_noise(43, nblk * B)
    drv = OracleDriver(B, 3, 1, max_blocks=1, max_delay=64, fractional_delay=True)
    R = drv.ring_length
    f1, f2 = drv.filter(h1), drv.filter(h2)
    d1, d2 = 16.0 + 37.3 * 3 / 11, 23.71
    drv.select(0, f1, delay=d1)
    outs = []
    for i in range(nblk):
        if i == sw:
            drv.select(0, f2, delay=d2, crossfade=True)
        outs.append(drv.process(x[i * B:(i + 1) * B], cl.FMT_FLOAT, 1, cl.FMT_FLOAT, 1, B).view(np.float32))
    y = np.concatenate(outs)
    # truth: stream = direct conv (crossfaded at block sw), then the polyphase read on a float64 ring
    y1, y2 = orc.direct(x, h1), orc.direct(x, h2)
    g = np.arange(B) / B
    s = y1.copy()
    s[sw * B:(sw + 1) * B] = (1 - g) * y1[sw * B:(sw + 1) * B] + g * y2[sw * B:(sw + 1) * B]
    s[(sw + 1) * B:] = y2[(sw + 1) * B:]
    want = np.zeros(nblk * B)
    ring = np.zeros(R)
    for i in range(nblk):
        w = (i * B) % R
        ring[(w + np.arange(B)) % R] = s[i * B:(i + 1) * B]
        n = np.arange(B)
        def rd(d):
            pos = np.fmod((w + n + R) - d, R)
            return orc.frac(ring, 0, 1, R, pos)
        if i < sw:
            want[i * B:(i + 1) * B] = rd(d1)
        elif i == sw:
            want[i * B:(i + 1) * B] = (1 - g) * rd(d1) + g * rd(d2)
        else:
            want[i * B:(i + 1) * B] = rd(d2)
    assert_float_parity(y, want, "fractional delay")


def test_convolver_routed_mixdown(orc):
    """Binaural shape: sources x 2 ears, time-domain MixSamples mixdown with gains, paths ascending."""
    B, L, nsrc, nblk = 64, 128, 5, 8
    xs = [make_noise(50 + s, nblk * B) for s in range(nsrc)]
    drv = OracleDriver(B, 2, nsrc, n_outputs=2, n_paths=2 * nsrc, mode=cl.MODE_ROUTED, max_blocks=2, max_delay=8)
    irs, gains = {}, {}
    for s in range(nsrc):
        for ear in range(2):
            p = 2 * s + ear
            irs[p] = make_ir(60 + p, L)
            gains[p] = 0.25 + 0.1 * p
            drv.route(p, s, ear, gains[p])
            drv.select(p, drv.filter(irs[p]), delay=float(ear * 3))
    y = run_float(drv, interleave(xs), 2 * B)
    for ear in range(2):
        want = np.zeros(nblk * B)
        for s in range(nsrc):
            p = 2 * s + ear
            yp = np.concatenate([np.zeros(ear * 3), orc.direct(xs[s], irs[p])])[: nblk * B]
            want += gains[p] * yp
        assert_float_parity(y[:, ear], want, "ear %d" % ear)


def test_convolver_mimo(orc):
    B, L, nin, nout, nblk = 64, 200, 3, 2, 8
    xs = [make_noise(70 + i, nblk * B) for i in range(nin)]
    drv = OracleDriver(B, 4, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=4)
    irs = {}
    for o in range(nout):
        for i in range(nin):
            irs[(o, i)] = make_ir(80 + o * nin + i, L)
            drv.select(o * nin + i, drv.filter(irs[(o, i)]))
    y = run_float(drv, interleave(xs), 4 * B)
    for o in range(nout):
        want = sum(orc.direct(xs[i], irs[(o, i)]) for i in range(nin))
        assert_float_parity(y[:, o], want, "out %d" % o)


def test_convolver_int24_io(orc):
    """s24 in / s24 out: inputs quantised through the float->s24 law, outputs within 1 LSB(24) of float64 truth
    (|error| <= 2^-23 from truncation plus the float tolerance)."""
    B, L, nch, nblk = 64, 100, 2, 6
    irs = [make_ir(90 + c, L) * 0.5 for c in range(nch)]
    xs = [make_noise(95 + c, nblk * B) * 0.9 for c in range(nch)]
    xi = interleave(xs)
    pcm = np.zeros(xi.size * 3, dtype=np.uint8)
    orc.transfer(xi.view(np.uint8).reshape(-1), cl.FMT_FLOAT, 0, 0, nch, pcm, cl.FMT_24, 0, 0, nch, nch, nblk * B)
    xq = s24_to_float(pcm).reshape(-1, nch)
    drv = OracleDriver(B, 2, nch, max_blocks=3)
    for c in range(nch):
        drv.select(c, drv.filter(irs[c]))
    out = np.concatenate([drv.process(pcm[i * 3 * B * nch * 3:(i + 1) * 3 * B * nch * 3], cl.FMT_24, nch, cl.FMT_24, nch, 3 * B)
                          for i in range(nblk // 3)])
    y = s24_to_float(out).reshape(-1, nch)
    for c in range(nch):
        want = orc.direct(xq[:, c], irs[c])
        assert np.abs(y[:, c] - want).max() <= 2.0 ** -23 + 1e-6
