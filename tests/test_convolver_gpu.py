"""GPU tier (-m gpu): the CUDA engine, called through the C ABI, against the CPU oracle on the same seeded
inputs (tolerance class: SNR >= 110 dB and max-abs <= 1e-5 x peak, BASELINE.json north_star), plus
size-independent properties at BASELINE.json's full sizes.  Mirrors tests/test_convolver_oracle.py.
"""
import numpy as np
import pytest

import cpulibs as cl
from convkit import GpuDriver, OracleDriver, interleave, make_ir, make_noise, run_float
from parity import assert_float_parity, compare_float, s24_to_float

pytestmark = pytest.mark.gpu


def both(bbx, *a, **kw):
    return GpuDriver(bbx, *a, **kw), OracleDriver(*a, **kw)


@pytest.mark.parametrize("B", [64, 128, 256, 512, 1024, 2048, 4096])
def test_blockconvolver_all_block_sizes(bbx, orc, B):
    """BlockConvolver::Convolve block by block, every supported partition size (T = 1)."""
    L, nblk = 3 * B + 17, 7
    h, x = make_ir(100 + B, L), make_noise(200 + B, nblk * B)
    bc = bbx.BlockConvolver(B, 4)
    bc.SetFilter(bc.CreateFilter(h))
    y = np.concatenate([bc.Convolve(x[i * B:(i + 1) * B]) for i in range(nblk)])
    bc.close()
    f = orc.filter(h, B)
    obc = orc.blockconv(B, 4)
    obc.set_filter(f)
    yo = np.concatenate([obc.convolve(x[i * B:(i + 1) * B]) for i in range(nblk)])
    assert_float_parity(y, yo, "vs oracle")
    assert_float_parity(y, orc.direct(x, h), "vs float64 direct")


@pytest.mark.parametrize("name,B,L,nch,nblk", [("C1", 1024, 8192, 2, 24), ("C2-path", 256, 512, 8, 16), ("C4", 512, 4096, 4, 20)])
def test_config_shapes_vs_oracle(bbx, orc, name, B, L, nch, nblk):
    P = -(-L // B)
    g, o = both(bbx, B, P, nch, max_blocks=8)
    irs = [make_ir(2000 + c, L) for c in range(nch)]
    xs = interleave([make_noise(1000 + c, nblk * B) for c in range(nch)])
    for c in range(nch):
        g.select(c, g.filter(irs[c]))
        o.select(c, o.filter(irs[c]))
    yg = run_float(g, xs, [B, 8 * B, 3 * B])  # ragged call sizes exercise the T-batched path and the state carry
    yo = run_float(o, xs, [B, 8 * B, 3 * B])
    g.close()
    for c in range(nch):
        assert_float_parity(yg[:, c], yo[:, c], "%s ch %d" % (name, c))


def test_long_reverb_shape_vs_oracle_and_direct(bbx, orc):
    """C3 shape at reduced channel count: 144000 taps, B = 512, P = 282, 6 channels, T up to 16."""
    B, L, nch, nblk = 512, 144000, 6, 300
    g, o = both(bbx, B, 282, nch, max_blocks=16)
    irs = [make_ir(2000 + c, L) for c in range(nch)]
    xs = [make_noise(1000 + c, nblk * B) for c in range(nch)]
    for c in range(nch):
        g.select(c, g.filter(irs[c]))
        o.select(c, o.filter(irs[c]))
    xi = interleave(xs)
    yg = run_float(g, xi, [16 * B, B, 7 * B])
    yo = run_float(o, xi, [16 * B, B, 7 * B])
    g.close()
    for c in range(nch):
        assert_float_parity(yg[:, c], yo[:, c], "C3 ch %d vs oracle" % c)
    n0 = nblk * B - 1024
    assert_float_parity(yg[n0:, 2], orc.direct(xs[2], irs[2], n0=n0, count=1024), "C3 vs float64 direct window")


def test_batching_invariance_bit_exact(bbx):
    """T blocks in one call == the same blocks one call at a time, bit for bit (block/partition indexing)."""
    B, L, nch, nblk = 256, 1500, 3, 12
    irs = [make_ir(300 + c, L) for c in range(nch)]
    xi = interleave([make_noise(310 + c, nblk * B) for c in range(nch)])
    outs = []
    for sizes in (B, 4 * B, [3 * B, B, 2 * B]):
        g = GpuDriver(bbx, B, 6, nch, max_blocks=4, max_delay=40, fractional_delay=True)
        for c in range(nch):
            g.select(c, g.filter(irs[c]), delay=5.3 * c)
        outs.append(run_float(g, xi, sizes))
        g.close()
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert np.array_equal(outs[0].view(np.uint32), outs[2].view(np.uint32))


@pytest.mark.parametrize("kw", [dict(mac_time_tile=1), dict(mac_time_tile=32), dict(mac_time_tile=116), dict(mac_time_tile=1, mac_l2_keep_16ths=5),
                                dict(mac_time_tile=1, mac_ctas_per_sm=2)])
def test_mac_variants_bit_identical(bbx, kw):
    """Every MAC kernel variant (streaming; time-batched: the shared-stream kernel k_fdl_mac_tbs (default) and round 1's
    per-thread-copy kernels, tile 116 = TT 16, 32 = TT 32; L2-residency hints; other unroll depth) uses the same plan and
    the same per-output FMA order: outputs must be bit-identical to the default, including across a crossfaded filter
    switch, ragged call sizes and a MIMO matrix."""
    B, L, nch, nblk = 512, 20000, 5, 76
    irs = [make_ir(400 + c, L) for c in range(nch + 1)]
    xi = interleave([make_noise(410 + c, nblk * B) for c in range(nch)])
    outs = []
    for extra in ({}, kw):
        g = GpuDriver(bbx, B, 40, nch, max_blocks=40, max_delay=30, fractional_delay=True, **extra)
        fl = [g.filter(h) for h in irs]
        for c in range(nch):
            g.select(c, fl[c], delay=2.5 * c)
        a = run_float(g, xi[:46 * B], [40 * B, 6 * B])
        g.select(2, fl[nch], delay=11.25, crossfade=True)
        b = run_float(g, xi[46 * B:], [23 * B, 7 * B])
        outs.append(np.concatenate([a, b]))
        g.close()
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    # MIMO: several terms per job accumulate in registers across segments
    nin, nout, Lm = 6, 3, 3000
    xm = interleave([make_noise(500 + i, 36 * B) for i in range(nin)])
    outs = []
    for extra in ({}, kw):
        g = GpuDriver(bbx, B, 6, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=36, **extra)
        for o in range(nout):
            for i in range(nin):
                g.select(o * nin + i, g.filter(make_ir(600 + o * nin + i, Lm)))
        outs.append(run_float(g, xm, 36 * B))
        g.close()
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


@pytest.mark.parametrize("B,P,nch,calls,occ", [
    (64, 40, 3, [64, 17, 33, 100, 9, 64], 0),      # B = 64: four tiles on 64 columns; more than one tile group (100 blocks)
    (128, 37, 5, [20, 32, 17, 64, 31], 0),          # B = 128: two tiles (17..32 blocks) and four
    (512, 45, 20, [64, 48, 33, 64], 2),             # 45 partitions (13 mod 16), 296 row ranges: many short segments per CTA
    (256, 70, 9, [64, 64, 24, 64, 64], 0),          # FDL ring wraps inside the calls (R = 70 + 64 - 1)
    (512, 36, 150, [64, 40], 0),                    # many channels: several whole filters per row range
])
def test_shared_stream_mac_shapes_bit_identical(bbx, B, P, nch, calls, occ):
    """The shared-stream time-batched kernels -- k_fdl_mac_tbw (warp-private rings, four tiles, calls of more than 32 blocks)
    and k_fdl_mac_tbs<2> (block-shared rings, 17..32 blocks) -- against the streaming kernel on the same plan: bit-identical
    for every tile count, partial tile groups, ragged segment lengths, ring wrap, and row ranges that hold several segments."""
    L = P * B - B // 3
    irs = [make_ir(900 + c, L - 7 * (c % 5) * B // 8) for c in range(nch)]  # a few shorter filters: uneven terms
    total = sum(calls)
    xi = interleave([make_noise(950 + c, total * B) for c in range(nch)])
    outs = []
    for tile in (0, 1):
        g = GpuDriver(bbx, B, P, nch, max_blocks=max(calls), mac_time_tile=tile, mac_ctas_per_sm=occ)
        for c in range(nch):
            g.select(c, g.filter(irs[c]))
        outs.append(run_float(g, xi, [n * B for n in calls]))
        name = g.eng.mac_kernel_name()
        g.close()
        assert name.startswith(("k_fdl_mac_tbw", "k_fdl_mac_tbs")) if tile == 0 else name == "k_fdl_mac", name
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def test_shared_stream_mac_vs_oracle(bbx, orc):
    """the default batched path against the CPU oracle and float64 direct convolution (not only against its sibling kernel)"""
    B, L, nch, nblk = 256, 256 * 36 + 100, 4, 96
    g, o = both(bbx, B, 37, nch, max_blocks=64)
    irs = [make_ir(970 + c, L) for c in range(nch)]
    xs = [make_noise(980 + c, nblk * B) for c in range(nch)]
    for c in range(nch):
        g.select(c, g.filter(irs[c]))
        o.select(c, o.filter(irs[c]))
    xi = interleave(xs)
    yg = run_float(g, xi, [64 * B, 32 * B])
    yo = run_float(o, xi, [64 * B, 32 * B])
    assert g.eng.mac_kernel_name().startswith("k_fdl_mac_tbs<2")  # the last call has 32 blocks: two tiles per CTA
    g.close()
    for c in range(nch):
        assert_float_parity(yg[:, c], yo[:, c], "vs oracle ch %d" % c)
    assert_float_parity(yg[:, 1], orc.direct(xs[1], irs[1]), "vs float64 direct")


@pytest.mark.parametrize("direct", [True, False])
def test_async_host_pipeline_matches_sync(bbx, direct):
    """bbx_process_async with pinned I/O == the synchronous call, bit for bit: short calls through the direct path
    (kernels read / write the pinned buffers over PCIe) and through the staged copy-engine pipeline."""
    B, L, nch, nblk, T = 256, 3000, 32, 24, 4   # 32 float channels = 128 B per frame: wide enough for the direct path
    irs = [make_ir(700 + c, L) for c in range(nch)]
    xi = interleave([make_noise(710 + c, nblk * B) for c in range(nch)])
    g = GpuDriver(bbx, B, 12, nch, max_blocks=T)
    for c in range(nch):
        g.select(c, g.filter(irs[c]))
    want = run_float(g, xi, T * B)
    g.close()
    g = GpuDriver(bbx, B, 12, nch, max_blocks=T)
    for c in range(nch):
        g.select(c, g.filter(irs[c]))
    nbytes = T * B * nch * 4
    if not direct:
        g.eng.set_direct_io(0)
    hin = [bbx.PinnedBuffer(nbytes) for _ in range(nblk // T)]
    hout = [bbx.PinnedBuffer(nbytes) for _ in range(nblk // T)]
    for i in range(nblk // T):
        hin[i].array[:] = xi[i * T * B:(i + 1) * T * B].reshape(-1).view(np.uint8)
        g.eng.ConvolveHostPtrAsync(hin[i].ptr, cl.FMT_FLOAT, nch, hout[i].ptr, cl.FMT_FLOAT, nch, T * B)
    g.eng.Sync()
    assert g.eng.direct_calls() == (nblk // T if direct else 0)
    got = np.concatenate([h.array.view(np.float32).reshape(T * B, nch).copy() for h in hout])
    g.close()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_zero_input_exact_zeros_and_null_filter(bbx):
    B = 128
    g = GpuDriver(bbx, B, 4, 2, max_blocks=2)
    g.select(0, g.filter(make_ir(1, 500)))
    y = run_float(g, np.zeros((4 * B, 2), dtype=np.float32), 2 * B)
    assert not y.any()
    x = interleave([make_noise(2, 4 * B), make_noise(3, 4 * B)])
    y = run_float(g, x, 2 * B)
    assert y[:, 0].any() and not y[:, 1].any()  # path 1 has no filter: silence
    g.select(0, None)
    y = run_float(g, x, 2 * B)
    assert not y.any()
    g.close()


@pytest.mark.parametrize("B,L", [(64, 300), (512, 4096)])
def test_delayed_impulse_ir(bbx, B, L):
    P = -(-L // B)
    nblk = P + 8
    x = make_noise(1, nblk * B)
    for tap in (0, B - 1, B, B + 1, L - 1):
        h = np.zeros(L, dtype=np.float32)
        h[tap] = 1.0
        g = GpuDriver(bbx, B, P, 1, max_blocks=4)
        g.select(0, g.filter(h))
        y = run_float(g, x.reshape(-1, 1), 4 * B)[:, 0]
        g.close()
        want = np.concatenate([np.zeros(tap, dtype=np.float32), x])[: x.size]
        assert_float_parity(y, want, "tap %d" % tap)


def test_filter_switching_vs_oracle(bbx, orc):
    """C4 shape: bank of IRs, a switch every few blocks, crossfaded and hard, fractional delays, s24 in/out."""
    B, L, nch, nbank = 512, 4096, 4, 5
    g, o = both(bbx, B, 8, nch, max_blocks=10, max_delay=64, fractional_delay=True)
    assert g.ring_length == o.ring_length
    bank = [[make_ir(2000 + 16 * c + k, L) for k in range(nbank)] for c in range(nch)]
    gf = [[g.filter(h) for h in row] for row in bank]
    of = [[o.filter(h) for h in row] for row in bank]
    rng = np.random.default_rng(77)
    pcm_g, pcm_o = [], []
    for m in range(8):
        nblk = 9 if m % 2 else 10  # 100 ms at 48 kHz = every 9th/10th block of 512
        for c in range(nch):
            k = (m + c) % nbank
            d = 16 + 37.3 * ((m * 7 + c) % 11) / 11
            xf = (m % 3 != 2)
            if m == 0 or c != 3 or m % 2 == 0:  # channel 3 only switches on even m
                g.select(c, gf[c][k], delay=d, crossfade=xf and m > 0)
                o.select(c, of[c][k], delay=d, crossfade=xf and m > 0)
        x = rng.uniform(-0.25, 0.25, (nblk * B, nch)).astype(np.float32)  # keeps the s24 outputs off the clip rails
        pcm = np.zeros(x.size * 3, dtype=np.uint8)
        orc.transfer(x.view(np.uint8).reshape(-1), cl.FMT_FLOAT, 0, 0, nch, pcm, cl.FMT_24, 0, 0, nch, nch, nblk * B)
        pcm_g.append(g.process(pcm, cl.FMT_24, nch, cl.FMT_24, nch, nblk * B))
        pcm_o.append(o.process(pcm, cl.FMT_24, nch, cl.FMT_24, nch, nblk * B))
    g.close()
    yg = s24_to_float(np.concatenate(pcm_g)).reshape(-1, nch)
    yo = s24_to_float(np.concatenate(pcm_o)).reshape(-1, nch)
    for c in range(nch):
        r = compare_float(yg[:, c], yo[:, c])
        # int24 outputs (SURVEY.md 8.A parity classes): equal as floats within the float tolerance, plus one
        # LSB(24) where the float inputs to the converter straddle a quantisation step
        assert r["peak"] < 0.999, "test signal clipped"
        assert r["snr_db"] >= 110.0, r
        assert r["max_abs"] <= 2.0 ** -23 + 1e-5 * r["peak"], r


def test_delay_only_crossfade_and_integer_mode(bbx):
    B, L, nblk = 128, 400, 9
    for frac in (False, True):
        g, o = both(bbx, B, 4, 2, max_blocks=3, max_delay=50, fractional_delay=frac)
        hs = [make_ir(5, L), make_ir(6, L)]
        gf, of = [g.filter(h) for h in hs], [o.filter(h) for h in hs]
        xi = interleave([make_noise(7, nblk * B), make_noise(8, nblk * B)])
        ys = []
        for d in (g, o):
            fl = gf if d is g else of
            d.select(0, fl[0], delay=3)
            d.select(1, fl[1], delay=0)
            a = run_float(d, xi[:3 * B], 3 * B)
            d.select(0, fl[0], delay=41 if not frac else 40.6, crossfade=True)  # same filter, new delay
            d.select(1, fl[1], delay=50, crossfade=False)                       # hard delay jump
            b = run_float(d, xi[3 * B:], 3 * B)
            ys.append(np.concatenate([a, b]))
        g.close()
        for c in range(2):
            assert_float_parity(ys[0][:, c], ys[1][:, c], "delay switch frac=%s ch %d" % (frac, c))


def test_routed_binaural_vs_oracle(bbx):
    """C2 shape: 64 sources x 2 ears, 512-tap HRIRs, 256-sample blocks, per-path ITD delay and gain."""
    B, L, nsrc, nblk = 256, 512, 64, 12
    kw = dict(n_outputs=2, n_paths=2 * nsrc, mode=cl.MODE_ROUTED, max_blocks=4, max_delay=40)
    g, o = both(bbx, B, 2, nsrc, **kw)
    xi = interleave([make_noise(1000 + s, nblk * B) for s in range(nsrc)])
    for s in range(nsrc):
        for ear in range(2):
            p = 2 * s + ear
            h = make_ir(2000 + p, L)
            gain = 0.05 + 0.01 * (p % 7)
            delay = float((s * (1 + ear)) % 37)
            for d in (g, o):
                d.route(p, s, ear, gain)
                d.select(p, d.filter(h), delay=delay)
    yg, yo = run_float(g, xi, [4 * B, B]), run_float(o, xi, [4 * B, B])
    g.close()
    for ear in range(2):
        assert_float_parity(yg[:, ear], yo[:, ear], "ear %d" % ear)


@pytest.mark.parametrize("fractional,fmt", [(False, cl.FMT_FLOAT), (True, cl.FMT_FLOAT), (False, cl.FMT_24)])
def test_mixdown_kernels_bit_identical(bbx, fractional, fmt):
    """k_pcm_out_mix (route-parallel products, ordered sums) against the per-output kernel: same bytes for a many-path
    mixdown with zero gains, per-path delays, a delay crossfade on a switching call, ragged call sizes and three outputs."""
    B, L, nsrc, nout, nblk = 128, 200, 40, 3, 14
    xi = interleave([make_noise(1500 + s, nblk * B) for s in range(nsrc)])
    outs = []
    for per_output in (False, True):
        g = GpuDriver(bbx, B, 2, nsrc, n_outputs=nout, n_paths=2 * nsrc, mode=cl.MODE_ROUTED, max_blocks=4, max_delay=90,
                      fractional_delay=fractional)
        g.eng.set_mixdown_kernel(per_output)
        fl = [g.filter(make_ir(1600 + p, L)) for p in range(2 * nsrc)]
        for p in range(2 * nsrc):
            g.route(p, p % nsrc, (p * 7) % nout, 0.0 if p % 9 == 4 else 0.05 + 0.01 * (p % 5))
            g.select(p, fl[p], delay=(p % 13) * 6.25 if fractional else float((p * 5) % 80))
        got, pos = [], 0
        for k, nb in enumerate([4, 1, 3, 2, 4]):
            if k == 2:
                for p in (0, 5, 41):
                    g.select(p, fl[(p + 1) % (2 * nsrc)], delay=33.5 if fractional else 17.0, crossfade=True)
            y = g.process(xi[pos * B:(pos + nb) * B], cl.FMT_FLOAT, nsrc, fmt, nout + 1, nb * B)
            got.append(np.array(y, copy=True))
            pos += nb
        outs.append(np.concatenate(got))
        g.close()
    assert outs[0].any() and np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("fmt", [cl.FMT_FLOAT, cl.FMT_24, cl.FMT_16])
def test_fractional_output_kernels_bit_identical(bbx, fmt):
    """k_pcm_out_frac (one sample per thread, a warp per frame) against the tile kernel: same bytes for per-channel
    fractional delays, a crossfaded delay + filter switch, a hard switch, 35 channels (a ragged channel tile) and an
    unused PCM channel."""
    B, L, nch, nblk = 64, 150, 35, 12
    xi = interleave([make_noise(1700 + c, nblk * B) for c in range(nch)]) * 0.5
    outs = []
    for tile_kernel in (False, True):
        g = GpuDriver(bbx, B, 3, nch, max_blocks=5, max_delay=70, fractional_delay=True)
        g.eng.set_mixdown_kernel(tile_kernel)
        fl = [g.filter(make_ir(1800 + c, L)) for c in range(nch + 2)]
        for c in range(nch):
            g.select(c, fl[c], delay=(c % 9) * 7.3)
        got, pos = [], 0
        for k, nb in enumerate([5, 1, 3, 3]):
            if k == 1:
                g.select(2, fl[nch], delay=55.125, crossfade=True)
                g.select(30, fl[30], delay=0.5, crossfade=True)
            if k == 3:
                g.select(7, fl[nch + 1], delay=12.75, crossfade=False)
            y = g.process(xi[pos * B:(pos + nb) * B], cl.FMT_FLOAT, nch, fmt, nch + 1, nb * B)
            got.append(np.array(y, copy=True))
            pos += nb
        outs.append(np.concatenate(got))
        g.close()
    assert outs[0].any() and np.array_equal(outs[0], outs[1])


def test_select_filters_batch(bbx):
    """bbx_set_filters == a loop of bbx_set_filter (same bytes), and a bad entry latches nothing."""
    B, L, nch, nblk = 64, 150, 5, 6
    xi = interleave([make_noise(1900 + c, nblk * B) for c in range(nch)])
    outs = []
    for batch in (False, True):
        g = GpuDriver(bbx, B, 3, nch, max_blocks=3, max_delay=20)
        fl = [g.filter(make_ir(1950 + c, L)) for c in range(2 * nch)]
        if batch:
            g.eng.SelectFilters(range(nch), fl[:nch], delays=[float(c) for c in range(nch)])
        else:
            for c in range(nch):
                g.select(c, fl[c], delay=float(c))
        a = run_float(g, xi[:3 * B], 3 * B)
        if batch:
            with pytest.raises(bbx.BbxError):   # path 99 does not exist: nothing of this request may be latched
                g.eng.SelectFilters([0, 99], [fl[nch], fl[nch + 1]], delays=[1.0, 1.0], crossfade=[True, True])
            g.eng.SelectFilters([1, 3], [fl[nch + 1], fl[nch + 3]], delays=[7.0, 0.0], crossfade=[True, False])
        else:
            g.select(1, fl[nch + 1], delay=7.0, crossfade=True)
            g.select(3, fl[nch + 3], delay=0.0, crossfade=False)
        b = run_float(g, xi[3 * B:], 3 * B)
        outs.append(np.concatenate([a, b]))
        g.close()
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def test_mimo_vs_oracle(bbx):
    """C5 shape at reduced size: 8 x 8 matrix of 4096-tap IRs, B = 512, frequency-domain mixdown."""
    B, L, nin, nout, nblk = 512, 4096, 8, 8, 12
    g, o = both(bbx, B, 8, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=4)
    xi = interleave([make_noise(1000 + i, nblk * B) for i in range(nin)])
    for oo in range(nout):
        for i in range(nin):
            h = make_ir(2000 + 64 * oo + i, L)
            for d in (g, o):
                d.select(oo * nin + i, d.filter(h))
    ya = [run_float(d, xi[:8 * B], 4 * B) for d in (g, o)]
    # crossfade two entries of the matrix, hard-switch one
    h2 = [make_ir(5000 + k, L) for k in range(3)]
    for d in (g, o):
        d.select(0 * nin + 1, d.filter(h2[0]), crossfade=True)
        d.select(3 * nin + 2, d.filter(h2[1]), crossfade=True)
        d.select(5 * nin + 5, d.filter(h2[2]), crossfade=False)
    yb = [run_float(d, xi[8 * B:], 4 * B) for d in (g, o)]
    g.close()
    yg, yo = np.concatenate([ya[0], yb[0]]), np.concatenate([ya[1], yb[1]])
    for oo in range(nout):
        assert_float_parity(yg[:, oo], yo[:, oo], "MIMO out %d" % oo)


def test_formats_in_out_and_extra_channels(bbx, orc):
    """Every PCM format both ways (LE and BE); channels beyond n_inputs / n_outputs are ignored / preserved."""
    B, L, nch, nblk = 64, 100, 2, 4
    irs = [make_ir(90 + c, L) * 0.5 for c in range(nch)]
    x = (interleave([make_noise(95 + c, nblk * B) for c in range(3)]) * 0.9).astype(np.float32)  # 3 in-channels
    for fmt in (cl.FMT_16, cl.FMT_24, cl.FMT_32, cl.FMT_FLOAT, cl.FMT_DOUBLE):
        for be in (False, True):
            pcm = np.zeros(x.size * cl.FMT_BYTES[fmt], dtype=np.uint8)
            orc.transfer(x.view(np.uint8).reshape(-1), cl.FMT_FLOAT, 0, 0, 3, pcm, fmt, be, 0, 3, 3, nblk * B)
            outs = []
            for mk in (lambda: GpuDriver(bbx, B, 2, nch, max_blocks=4), lambda: OracleDriver(B, 2, nch, max_blocks=4)):
                d = mk()
                for c in range(nch):
                    d.select(c, d.filter(irs[c]))
                out = np.full(nblk * B * 4 * cl.FMT_BYTES[fmt], 0x3C, dtype=np.uint8)  # 4 out-channels
                if d.name == "gpu":
                    d.eng.Convolve(pcm, fmt, 3, fmt, 4, nblk * B, be, be, out=out)
                else:
                    d.cv.process(pcm, fmt, 3, fmt, 4, nblk * B, be, be, out=out)
                outs.append(out)
                d.close()
            a = outs[0].reshape(nblk * B, 4, -1)
            b = outs[1].reshape(nblk * B, 4, -1)
            assert (a[:, 2:] == 0x3C).all() and (b[:, 2:] == 0x3C).all()  # untouched channels
            fa = np.zeros(nblk * B * 2, dtype=np.float64)
            fb = np.zeros(nblk * B * 2, dtype=np.float64)
            for arr, dst in ((a, fa), (b, fb)):
                src = np.ascontiguousarray(arr[:, :2]).reshape(-1)
                orc.transfer(src, fmt, be, 0, 2, dst.view(np.uint8), cl.FMT_DOUBLE, 0, 0, 2, 2, nblk * B)
            lsb = {cl.FMT_16: 2.0 ** -15, cl.FMT_24: 2.0 ** -23}.get(fmt, 0.0)
            r = compare_float(fa, fb)
            assert r["max_abs"] <= lsb + 1e-5 * r["peak"], (fmt, be, r)


@pytest.mark.parametrize("fmt,nch,routed,expect_direct", [
    (cl.FMT_FLOAT, 32, False, True),    # 128 B of used channels per frame on both sides: both direct
    (cl.FMT_16, 64, False, True),
    (cl.FMT_16, 32, False, False),      # 64 B per frame: too narrow for the bus
    (cl.FMT_24, 48, False, False),      # 3-byte samples are byte-wise accesses: staged
    (cl.FMT_FLOAT, 32, True, True),     # wide input direct, 2-channel mixdown output staged
])
def test_direct_io_formats_and_extra_channels(bbx, orc, fmt, nch, routed, expect_direct):
    """The latency path (pinned buffers: the PCM kernels address host memory over PCIe, decided per side) gives the
    same bytes as the staged path and as pageable buffers, and leaves the channels beyond n_outputs untouched."""
    B, L, nblk = 128, 300, 5
    nout = 2 if routed else nch
    irs = [make_ir(190 + c, L) * 0.25 for c in range(nch)]
    x = (interleave([make_noise(195 + c, nblk * B) for c in range(nch + 1)]) * 0.9).astype(np.float32)  # one extra in-channel
    bps = cl.FMT_BYTES[fmt]
    pcm = np.zeros(x.size * bps, dtype=np.uint8)
    orc.transfer(x.view(np.uint8).reshape(-1), cl.FMT_FLOAT, 0, 0, nch + 1, pcm, fmt, 0, 0, nch + 1, nch + 1, nblk * B)
    res = {}
    for mode in ("direct", "staged", "pageable"):
        if routed:
            g = GpuDriver(bbx, B, 3, nch, n_outputs=2, n_paths=nch, mode=cl.MODE_ROUTED, max_blocks=1, max_delay=20)
            for c in range(nch):
                g.route(c, c, c % 2, 0.5)
        else:
            g = GpuDriver(bbx, B, 3, nch, max_blocks=1, max_delay=20)
        for c in range(nch):
            g.select(c, g.filter(irs[c]), delay=float(c % 7))
        if mode == "staged":
            g.eng.set_direct_io(0)
        nin, nob = B * (nch + 1) * bps, B * (nout + 2) * bps  # two extra out-channels
        if mode == "pageable":
            hin, hout = np.zeros(nin, dtype=np.uint8), np.zeros(nob, dtype=np.uint8)
            ain, aout = hin.ctypes.data, hout.ctypes.data
        else:
            pin, pout = bbx.PinnedBuffer(nin), bbx.PinnedBuffer(nob)
            hin, hout = pin.array, pout.array
            ain, aout = pin.ptr, pout.ptr
        got = []
        for b in range(nblk):
            hin[:] = pcm[b * nin:(b + 1) * nin]
            hout[:] = 0x3C
            g.eng.ConvolveHostPtr(ain, fmt, nch + 1, aout, fmt, nout + 2, B)
            got.append(hout.copy())
        assert g.eng.direct_calls() == (nblk if (mode == "direct" and expect_direct) else 0)
        g.close()
        res[mode] = np.concatenate(got).reshape(nblk * B, nout + 2, bps)
    assert (res["direct"][:, nout:] == 0x3C).all()
    assert res["direct"][:, :nout].any()
    assert np.array_equal(res["direct"], res["staged"]) and np.array_equal(res["direct"], res["pageable"])


def test_engine_argument_errors(bbx):
    with pytest.raises(bbx.BbxError):
        bbx.Convolver(100, 4, 2)  # not a power of two
    eng = bbx.Convolver(64, 2, 2, max_blocks=2, max_delay=10)
    with pytest.raises(bbx.BbxError):
        eng.CreateFilter(np.zeros(64 * 3, dtype=np.float32))  # 3 partitions > max_partitions
    f = eng.CreateFilter(np.ones(10, dtype=np.float32))
    with pytest.raises(bbx.BbxError):
        eng.SelectFilter(5, f)  # path out of range
    with pytest.raises(bbx.BbxError):
        eng.SelectFilter(0, f, delay=11.0)  # beyond max_delay
    x = np.zeros((64 * 3, 2), dtype=np.float32)
    with pytest.raises(bbx.BbxError):
        eng.Convolve(x, cl.FMT_FLOAT, 2, cl.FMT_FLOAT, 2, 64 * 3)  # 3 blocks > max_blocks
    with pytest.raises(bbx.BbxError):
        eng.Convolve(x[:100], cl.FMT_FLOAT, 2, cl.FMT_FLOAT, 2, 100)  # not a multiple of B
    eng.close()


# ---- BASELINE.json full sizes: size-independent properties -------------------------------------------
def test_full_size_c3_properties(bbx, orc):
    """128 channels x 144000 taps, B = 512 (P = 282, ~300 MB of spectra + FDL):
      - channel c convolved with a pure delay of c samples (impulse IR at tap 140000 + c) returns the input
        delayed (no oracle needed),
      - linearity: process(a x1 + b x2) == a process(x1) + b process(x2),
      - channel-shard invariance: channels 16..31 alone (the 8-GPU shard of rank 1) are bit-identical,
      - two spot channels with noise IRs against the float64 direct convolution on a window."""
    B, L, nch, nblk, T = 512, 144000, 128, 320, 32
    eng = GpuDriver(bbx, B, 282, nch, max_blocks=T)
    taps = [140000 + c for c in range(nch)]
    irs = {}
    for c in range(nch):
        if c in (5, 77):
            irs[c] = make_ir(2000 + c, L)
        else:
            h = np.zeros(L, dtype=np.float32)
            h[taps[c]] = 1.0
            irs[c] = h
        eng.select(c, eng.filter(irs[c]))
    x1 = interleave([make_noise(1000 + c, nblk * B) for c in range(nch)])
    y1 = run_float(eng, x1, T * B)
    for c in range(0, nch, 9):
        if c in (5, 77):
            continue
        want = np.concatenate([np.zeros(taps[c], dtype=np.float32), x1[:, c]])[: nblk * B]
        assert_float_parity(y1[:, c], want, "pure delay ch %d" % c)
    n0 = nblk * B - 512
    for c in (5, 77):
        assert_float_parity(y1[n0:, c], orc.direct(x1[:, c], irs[c], n0=n0, count=512), "noise IR ch %d" % c)
    eng.close()

    # linearity on a fresh engine state
    x2 = interleave([make_noise(3000 + c, nblk * B) for c in range(nch)])
    xs = (0.5 * x1 + 0.25 * x2).astype(np.float32)
    outs = []
    for x in (x2, xs):
        e2 = GpuDriver(bbx, B, 282, nch, max_blocks=T)
        for c in range(nch):
            e2.select(c, e2.filter(irs[c]))
        outs.append(run_float(e2, x, T * B))
        e2.close()
    lin = 0.5 * y1.astype(np.float64) + 0.25 * outs[0].astype(np.float64)
    for c in (0, 5, 64, 77, 127):
        assert_float_parity(outs[1][:, c], lin[:, c], "linearity ch %d" % c)

    # shard invariance: rank 1 of 8 owns channels 16..31
    first, count = bbx.shard_range(nch, 1, 8)
    assert (first, count) == (16, 16)
    e3 = GpuDriver(bbx, B, 282, count, max_blocks=T)
    for c in range(count):
        e3.select(c, e3.filter(irs[first + c]))
    ys = run_float(e3, np.ascontiguousarray(x1[:, first:first + count]), T * B)
    e3.close()
    assert np.array_equal(ys.view(np.uint32), np.ascontiguousarray(y1[:, first:first + count]).view(np.uint32))


def test_full_size_c5_mimo_property(bbx, orc):
    """64 x 64 matrix of 4096-tap IRs (134 MB of spectra): a matrix of scaled pure delays makes every output a
    known mix of delayed inputs; two noise-IR rows are checked against float64 direct sums on a window."""
    B, L, nin, nout, nblk, T = 512, 4096, 64, 64, 24, 8
    eng = GpuDriver(bbx, B, 8, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=T)
    xi = interleave([make_noise(1000 + i, nblk * B) for i in range(nin)])
    noise_rows = {3: {}, 40: {}}
    for o in range(nout):
        for i in range(nin):
            if o in noise_rows:
                h = make_ir(2000 + 64 * o + i, L) * 0.2
                noise_rows[o][i] = h
            else:
                h = np.zeros(L, dtype=np.float32)
                h[(37 * o + 11 * i) % L] = 1.0 / 64 if (o + i) % 2 == 0 else -1.0 / 64
            eng.select(o * nin + i, eng.filter(h))
    y = run_float(eng, xi, T * B)
    eng.close()
    for o in (0, 17, 63):
        want = np.zeros(nblk * B)
        for i in range(nin):
            d = (37 * o + 11 * i) % L
            s = 1.0 / 64 if (o + i) % 2 == 0 else -1.0 / 64
            want += s * np.concatenate([np.zeros(d), xi[:, i].astype(np.float64)])[: nblk * B]
        assert_float_parity(y[:, o], want, "MIMO delay matrix out %d" % o)
    n0 = nblk * B - 256
    for o, row in noise_rows.items():
        want = sum(orc.direct(xi[:, i], row[i], n0=n0, count=256) for i in range(nin))
        assert_float_parity(y[n0:, o], want, "MIMO noise row %d" % o)


@pytest.mark.parametrize("B,nch,L,T", [(64, 2000, 100, 8), (512, 600, 700, 4), (4096, 20, 5000, 8)])
def test_persistent_fft_many_items(bbx, orc, B, nch, L, T):
    """More (channel group, block) items than resident CTAs: every CTA of the persistent radix-8 FFT kernels
    (k_rfft8 / k_irfft8) loops over several items with its prefetch and cached twiddles; ragged last channel group."""
    P = -(-L // B)
    g = GpuDriver(bbx, B, P, nch, max_blocks=T)
    irs = [make_ir(900 + (c % 7), L) for c in range(7)]
    fl = [g.filter(h) for h in irs]
    for c in range(nch):
        g.select(c, fl[c % 7])
    rng = np.random.default_rng(77)
    x = rng.uniform(-1, 1, (2 * T * B, nch)).astype(np.float32)
    y = run_float(g, x, T * B)
    g.close()
    for c in (0, 1, nch // 2, nch - 2, nch - 1):
        f = orc.filter(irs[c % 7], B)
        obc = orc.blockconv(B, P)
        obc.set_filter(f)
        yo = np.concatenate([obc.convolve(np.ascontiguousarray(x[i * B:(i + 1) * B, c])) for i in range(2 * T)])
        assert_float_parity(y[:, c], yo, "channel %d of %d" % (c, nch))


@pytest.mark.parametrize("B", [64, 128, 256, 512, 1024, 2048, 4096])
def test_forward_fft_vs_float64_fft(bbx, B):
    """The forward transform kernels alone (a14: FFT): the filter object's spectra against numpy's float64 FFT of every
    zero-padded partition -- H[p] = R2C_2B([h[pB .. pB+B-1], 0^B]) / 2B, bin 0 packed as (DC, Nyquist) -- for every
    block size (radix-8 and radix-4 paths), a ragged last partition and a KAT (unit impulse -> flat spectrum 1/2B)."""
    rng = np.random.default_rng(4000 + B)
    L = 5 * B - 37
    h = rng.standard_normal(L).astype(np.float32)
    eng = bbx.Convolver(B, 6, 1)
    f = eng.CreateFilter(h)
    got = f.Spectra()
    P = -(-L // B)
    assert got.shape == (P, B)
    hp = np.zeros(P * B, dtype=np.float64)
    hp[:L] = h
    worst = 0.0
    for p in range(P):
        w = np.zeros(2 * B, dtype=np.float64)
        w[:B] = hp[p * B:(p + 1) * B]
        ref = np.fft.rfft(w) / (2 * B)
        want = ref[:B].copy()
        want[0] = ref[0].real + 1j * ref[B].real
        scale = np.abs(ref).max()
        worst = max(worst, np.abs(got[p].astype(np.complex128) - want).max() / scale)
    assert worst <= 2e-6, worst  # fp32 FFT of 2B points: a few ulp per pass, log2(2B) passes
    imp = np.zeros(B, dtype=np.float32)
    imp[0] = 1.0
    s = eng.CreateFilter(imp).Spectra()
    assert np.array_equal(s.real, np.full((1, B), 1.0 / (2 * B), dtype=np.float32))
    assert np.array_equal(s.imag[0, 1:], np.zeros(B - 1, dtype=np.float32)) and s.imag[0, 0] == np.float32(1.0 / (2 * B))
    eng.close()


def test_filter_destroy_refused_while_selected(bbx):
    """Ownership rule of include/bbx.h: a filter that a path still has selected (current or latched) cannot be destroyed --
    the MAC plans hold device pointers into its spectra; once another filter has taken over it can, and whatever is left
    dies with the engine (destroying such a handle later is a no-op)."""
    B = 128
    eng = bbx.Convolver(B, 4, 2, max_blocks=2)
    f0, f1, f2 = (eng.CreateFilter(make_ir(300 + k, 3 * B)) for k in range(3))
    x = interleave([make_noise(310 + c, 2 * B) for c in range(2)])
    eng.SelectFilter(0, f0)
    eng.SelectFilter(1, f1)
    with pytest.raises(bbx.BbxError, match="latched"):
        f0.close()  # latched, not yet applied
    y0 = eng.Convolve(x, bbx.FMT_FLOAT, 2, bbx.FMT_FLOAT, 2, 2 * B).view(np.float32).copy()
    with pytest.raises(bbx.BbxError, match="current"):
        f0.close()
    eng.SelectFilter(0, f2)
    eng.Convolve(x, bbx.FMT_FLOAT, 2, bbx.FMT_FLOAT, 2, 2 * B)
    f0.close()  # path 0 runs f2 now: f0 can go
    assert f0.h is None
    y1 = eng.Convolve(x, bbx.FMT_FLOAT, 2, bbx.FMT_FLOAT, 2, 2 * B).view(np.float32)
    assert np.isfinite(y1).all() and np.abs(y0).max() > 0
    h1 = f1.h
    eng.close()  # releases f1 and f2
    assert f1.h is None and f2.h is None
    assert bbx.lib().bbx_filter_destroy(h1) == 0  # a stale handle of a dead engine is ignored, not dereferenced


def test_rejected_call_consumes_nothing(bbx, orc):
    """bbx_process_async validates the geometry before it stages, copies or latches anything: a rejected call leaves the
    stream of results unchanged."""
    B, nch = 128, 3
    g, o = both(bbx, B, 4, nch, max_blocks=2)
    for c in range(nch):
        h = make_ir(400 + c, 2 * B + 5)
        g.select(c, g.filter(h))
        o.select(c, o.filter(h))
    xs = interleave([make_noise(410 + c, 6 * B) for c in range(nch)])
    yg = [g.process(xs[:2 * B], cl.FMT_FLOAT, nch, cl.FMT_FLOAT, nch, 2 * B).view(np.float32).reshape(-1, nch)]
    flat = xs.reshape(-1)
    for nframes, in_ch in ((B + 1, nch), (3 * B, nch), (2 * B, nch - 1)):  # not whole blocks / more than max_blocks / too few channels
        with pytest.raises(bbx.BbxError):
            g.eng.Convolve(flat[:nframes * in_ch], bbx.FMT_FLOAT, in_ch, bbx.FMT_FLOAT, nch, nframes)
    yg.append(g.process(xs[2 * B:4 * B], cl.FMT_FLOAT, nch, cl.FMT_FLOAT, nch, 2 * B).view(np.float32).reshape(-1, nch))
    yg.append(g.process(xs[4 * B:], cl.FMT_FLOAT, nch, cl.FMT_FLOAT, nch, 2 * B).view(np.float32).reshape(-1, nch))
    yo = run_float(o, xs, 2 * B)
    g.close()
    yg = np.concatenate(yg)
    for c in range(nch):
        assert_float_parity(yg[:, c], yo[:, c], "after rejected calls, ch %d" % c)


def test_objects_keep_their_device(bbx, gpu):
    """Handle-based entry points run on the device their object was created on and give the caller's device back: an
    engine on GPU 1 next to a delay buffer and a biquad bank created on GPU 0, used alternately from one thread."""
    if bbx.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch
    torch.cuda.set_device(0)
    d = bbx.SoundDelayBuffer()
    d.SetSize(2, 64, bbx.FMT_FLOAT)
    bq = bbx.BiQuadBank(2)
    bq.CalcCoeffs(bbx.BIQUAD_LPF12, 1000.0, 48000.0)
    B = 128
    eng = bbx.Convolver(B, 2, 2, max_blocks=1, device=1)
    h = make_ir(500, B + 9)
    for c in range(2):
        eng.SelectFilter(c, eng.CreateFilter(h))
    x = interleave([make_noise(510 + c, B) for c in range(2)])
    ref_bq = bbx.BiQuadBank(2)
    ref_bq.CalcCoeffs(bbx.BIQUAD_LPF12, 1000.0, 48000.0)
    for it in range(3):
        y = eng.Convolve(x, bbx.FMT_FLOAT, 2, bbx.FMT_FLOAT, 2, B).view(np.float32)
        assert torch.cuda.current_device() == 0, "the engine call must give the caller's device back"
        assert d.WriteSamples(x[:16], bbx.FMT_FLOAT, 0, 2, 16) == 16
        d.IncrementWritePosition(16)
        back = np.zeros(16 * 2, dtype=np.float32)
        assert d.ReadSamples(back, bbx.FMT_FLOAT, 16, 0, 2, 16) == 16
        assert np.array_equal(back.reshape(16, 2), x[:16])
        z = np.zeros_like(x)
        bq.Process(x, z, 2, 2, 2, B)
        assert np.isfinite(y).all() and np.isfinite(z).all()
    eng.close()
    d.close()
    bq.close()
    ref_bq.close()


@pytest.mark.gpu
def test_same_engine_on_two_devices_of_one_process(bbx, gpu):
    """Kernels that need more than the default dynamic shared memory (k_irfft8 with its staging area, the batched MAC) set
    their function attributes per device: the same engine created on GPU 0 and then on GPU 1 of one process gives the
    same bytes (block sizes of both radix-8 kinds, calls long enough for the multi-kernel and the time-batched paths)."""
    if bbx.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    for B, P, T in ((512, 3, 2), (64, 40, 40), (4096, 2, 2)):
        nch = 3
        irs = [make_ir(900 + c, P * B - 5) for c in range(nch)]
        x = interleave([make_noise(910 + c, T * B) for c in range(nch)])
        outs = []
        for dev in (0, 1):
            eng = bbx.Convolver(B, P, nch, max_blocks=T, device=dev)
            for c in range(nch):
                eng.SelectFilter(c, eng.CreateFilter(irs[c]))
            outs.append(np.concatenate([eng.Convolve(x, bbx.FMT_FLOAT, nch, bbx.FMT_FLOAT, nch, T * B) for _ in range(2)]))
            eng.close()
        assert np.array_equal(outs[0], outs[1]), "B = %d: GPU 1 differs from GPU 0" % B
        assert np.abs(outs[0].view(np.float32)).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["per_channel", "routed", "mimo"])
def test_state_checkpoint_resume(bbx, mode):
    """bbx_engine_get_state / set_state: a fresh engine with the same filters, resumed from a checkpoint taken in the middle
    of a crossfaded, delayed switch sequence, produces the bytes the original engine goes on to produce; so does the original
    engine rewound to the checkpoint."""
    B, L, nin = 128, 700, 3
    rng = np.random.default_rng(77)
    irs = [rng.standard_normal(L).astype(np.float32) * 0.05 for _ in range(6)]

    def make():
        if mode == "per_channel":
            e = bbx.Convolver(B, 8, nin, max_blocks=4, max_delay=300, fractional_delay=True)
            npaths, nout = nin, nin
        elif mode == "routed":
            npaths, nout = 5, 2
            e = bbx.Convolver(B, 8, nin, n_outputs=nout, n_paths=npaths, mode=bbx.MODE_ROUTED, max_blocks=4, max_delay=300)
            for k in range(npaths):
                e.SetRoute(k, k % nin, k % nout, 0.5 + 0.1 * k)
        else:
            nout = 2
            npaths = nin * nout
            e = bbx.Convolver(B, 8, nin, n_outputs=nout, mode=bbx.MODE_MIMO, max_blocks=4)
        fl = [e.CreateFilter(h) for h in irs]
        for k in range(npaths):
            e.SelectFilter(k, fl[k % len(fl)], delay=(0.0 if mode == "mimo" else 3.25 * k if mode == "per_channel" else 2.0 * k))
        return e, fl, npaths, nout

    def block(i, nblk):
        return np.random.default_rng(1000 + i).uniform(-1, 1, (nblk * B, nin)).astype(np.float32)

    def run(e, fl, npaths, nout, i, nblk):
        return e.Convolve(block(i, nblk), bbx.FMT_FLOAT, nin, bbx.FMT_FLOAT, nout, nblk * B).copy()

    sizes = [2, 1, 3, 1, 4, 2]
    a = make()
    for i in range(2):
        run(*a, i, sizes[i])
    # latch a crossfaded switch, THEN checkpoint: the pending selection is part of the state
    for k in range(a[2]):
        a[0].SelectFilter(k, a[1][(k + 2) % len(irs)], delay=(0.0 if mode == "mimo" else 7.5 + k if mode == "per_channel" else 5.0 + k),
                          crossfade=True)
    state = a[0].GetState()
    want = [run(*a, i, sizes[i]) for i in range(2, 6)]
    b = make()
    b[0].SetState(state)
    got = [run(*b, i, sizes[i]) for i in range(2, 6)]
    assert any(np.any(w) for w in want)
    for w, g in zip(want, got):
        assert np.array_equal(w, g)
    a[0].SetState(state)  # rewind the original
    again = [run(*a, i, sizes[i]) for i in range(2, 6)]
    for w, g in zip(want, again):
        assert np.array_equal(w, g)
    # a state of another geometry is refused and changes nothing
    c = bbx.Convolver(B, 8, nin + 1, max_blocks=4)
    with pytest.raises(bbx.BbxError):
        c.SetState(state)
    with pytest.raises(bbx.BbxError):
        b[0].SetState(state[:100])
    c.close()
    a[0].close()
    b[0].close()
