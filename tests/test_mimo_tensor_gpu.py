"""GPU tier (-m gpu): the MIMO per-bin complex GEMM on the tcgen05 tensor cores (k_mimo_tc, 3xTF32) against the
CPU oracle, float64 direct convolution and the SIMT MAC of the same engine.  Tolerance class of BASELINE.json's
north_star (SNR >= 110 dB, max-abs <= 1e-5 x peak): the tensor-core path is NOT bit-identical to the SIMT MAC.
"""
import numpy as np
import pytest

import cpulibs as cl
from convkit import GpuDriver, OracleDriver, interleave, make_ir, make_noise, run_float
from parity import assert_float_parity, compare_float

pytestmark = pytest.mark.gpu


def fill_matrix(drivers, nin, nout, L, seed0, null=(), lengths=None):
    for o in range(nout):
        for i in range(nin):
            if (o, i) in null:
                continue
            Li = lengths[(o * nin + i) % len(lengths)] if lengths else L
            h = make_ir(seed0 + 64 * o + i, Li)
            for d in drivers:
                d.select(o * nin + i, d.filter(h))


def tensor_ok(g, expect_launches=True):
    n, st = g.eng.tensor_status()
    assert st == 0, "k_mimo_tc reported a barrier time-out (status %d)" % st
    if expect_launches:
        assert n > 0, "the tensor-core path was not taken"
    return n


@pytest.mark.parametrize("B,nin,nout,L,calls", [
    (512, 8, 8, 4096, [32, 16, 20]),     # C5 shape at reduced size; N = 32, 16, 32 (ragged)
    (256, 5, 3, 1000, [64, 48, 17]),     # K = 5 x 4 = 20 complex -> padded to 32; n_out < 64 (padded rows)
    (64, 3, 70, 200, [40]),              # two output groups (70 outputs), P2 = 4, K = 12 -> 16
    (128, 2, 2, 128 * 16, [64, 64]),     # 16 partitions per filter, tiny matrix
    (256, 4, 4, 1024, [96, 80]),         # more than 64 blocks per call: two column tiles, the second one ragged
    (128, 3, 2, 100, [32]),              # single-partition filters (P2 = 1): 16 inputs' worth of K per chunk, padded
])
def test_mimo_tensor_vs_oracle(bbx, B, nin, nout, L, calls):
    P = -(-L // B)
    nblk = sum(calls)
    g = GpuDriver(bbx, B, P, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=max(calls))
    o = OracleDriver(B, P, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=max(calls))
    fill_matrix((g, o), nin, nout, L, 2000, null={(0, 1), (nout - 1, nin - 1)}, lengths=[L, max(1, L - B), max(1, L // 3)])
    xi = interleave([make_noise(1000 + i, nblk * B) for i in range(nin)])
    sizes = [c * B for c in calls]
    yg, yo = run_float(g, xi, sizes), run_float(o, xi, sizes)
    n = tensor_ok(g)
    assert n == len(calls)
    g.close()
    worst = min(assert_float_parity(yg[:, oo], yo[:, oo], "tensor MIMO out %d" % oo)["snr_db"] for oo in range(nout))
    print("B=%d %dx%d worst SNR vs oracle %.1f dB" % (B, nin, nout, worst))


def test_mimo_tensor_vs_simt_and_direct(bbx, orc):
    """Same engine, tensor cores on/off: both within tolerance of float64 direct convolution, and of each other."""
    B, L, nin, nout, nblk = 512, 4096, 16, 12, 48
    xi = interleave([make_noise(1100 + i, nblk * B) for i in range(nin)])
    ys = []
    for tensor_off in (0, 1):
        g = GpuDriver(bbx, B, 8, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=24, mimo_tensor=tensor_off)
        fill_matrix((g,), nin, nout, L, 3000)
        ys.append(run_float(g, xi, 24 * B))
        n = tensor_ok(g, expect_launches=not tensor_off)
        assert (n == 0) == bool(tensor_off)
        g.close()
    n0 = nblk * B - 512
    for oo in (0, 5, 11):
        assert_float_parity(ys[0][:, oo], ys[1][:, oo], "tensor vs SIMT out %d" % oo)
        want = sum(orc.direct(xi[:, i], make_ir(3000 + 64 * oo + i, L), n0=n0, count=512) for i in range(nin))
        for y, what in ((ys[0], "tensor"), (ys[1], "SIMT")):
            r = assert_float_parity(y[n0:, oo], want, "%s vs float64 direct out %d" % (what, oo))
            print(what, oo, r)


def test_mimo_tensor_filter_switches(bbx):
    """A crossfaded switch runs that call on the SIMT MAC (transitional plan), a hard switch repacks the operand;
    later calls are back on the tensor cores with the new matrix."""
    B, L, nin, nout, T = 256, 2048, 8, 8, 16
    g = GpuDriver(bbx, B, 8, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=T)
    o = OracleDriver(B, 8, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=T)
    fill_matrix((g, o), nin, nout, L, 4000)
    xi = interleave([make_noise(1200 + i, 5 * T * B) for i in range(nin)])
    ya = [run_float(d, xi[:2 * T * B], T * B) for d in (g, o)]
    assert tensor_ok(g) == 2
    h2 = [make_ir(5000 + k, L) for k in range(3)]
    for d in (g, o):
        d.select(0 * nin + 1, d.filter(h2[0]), crossfade=True)
        d.select(3 * nin + 2, d.filter(h2[1]), crossfade=True)
    yb = [run_float(d, xi[2 * T * B:3 * T * B], T * B) for d in (g, o)]
    assert tensor_ok(g) == 2  # the crossfade call used the SIMT plan
    for d in (g, o):
        d.select(5 * nin + 5, d.filter(h2[2]), crossfade=False)
        d.select(6 * nin + 0, None, crossfade=False)
    yc = [run_float(d, xi[3 * T * B:], T * B) for d in (g, o)]
    assert tensor_ok(g) == 4
    g.close()
    yg, yo = np.concatenate([ya[0], yb[0], yc[0]]), np.concatenate([ya[1], yb[1], yc[1]])
    for oo in range(nout):
        assert_float_parity(yg[:, oo], yo[:, oo], "MIMO switch out %d" % oo)


def test_mimo_tensor_batching_consistency(bbx):
    """64 blocks in one call (tensor cores) == the same blocks in calls of 4 (SIMT streaming MAC) within tolerance;
    zero input gives exact zeros on the tensor path."""
    B, L, nin, nout, nblk = 512, 4096, 8, 4, 64
    xi = interleave([make_noise(1300 + i, nblk * B) for i in range(nin)])
    outs = []
    for size in (64 * B, 4 * B):
        g = GpuDriver(bbx, B, 8, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=64)
        fill_matrix((g,), nin, nout, L, 6000)
        outs.append(run_float(g, xi, size))
        if size == 64 * B:
            z = run_float(g, np.zeros((2 * 64 * B, nin), dtype=np.float32), 64 * B)
            assert not z[64 * B:].any()  # the FDL history (8 partitions) has drained after one call of zeros
        g.close()
    for oo in range(nout):
        assert_float_parity(outs[0][:, oo], outs[1][:, oo], "batch 64 vs 4, out %d" % oo)


def test_full_size_c5_mimo_tensor_property(bbx, orc):
    """C5 at full size on the tensor cores: 64 x 64 matrix of 4096-tap IRs, 64-block calls.  A matrix of scaled
    pure delays makes every output a known mix of delayed inputs; two noise-IR rows against float64 direct sums."""
    B, L, nin, nout, nblk, T = 512, 4096, 64, 64, 128, 64
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
    assert tensor_ok(eng) == nblk // T
    eng.close()
    for o in (0, 17, 63):
        want = np.zeros(nblk * B)
        for i in range(nin):
            d = (37 * o + 11 * i) % L
            s = 1.0 / 64 if (o + i) % 2 == 0 else -1.0 / 64
            want += s * np.concatenate([np.zeros(d), xi[:, i].astype(np.float64)])[: nblk * B]
        r = assert_float_parity(y[:, o], want, "MIMO delay matrix out %d" % o)
        print("delay matrix", o, r)
    n0 = nblk * B - 256
    for o, row in noise_rows.items():
        want = sum(orc.direct(xi[:, i], row[i], n0=n0, count=256) for i in range(nin))
        r = assert_float_parity(y[n0:, o], want, "MIMO noise row %d" % o)
        print("noise row", o, r)


def test_mimo_tensor_longest_sum(bbx, orc):
    """K = 1024 complex terms per output bin (kTcMaxK, the longest sum the tensor-core path accepts: 128 inputs x 8
    partitions): accumulator truncation grows with K, the tolerance must still hold against float64 direct sums."""
    B, L, nin, nout, T = 128, 1024, 128, 2, 32
    g = GpuDriver(bbx, B, 8, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=T)
    irs = {}
    for o in range(nout):
        for i in range(nin):
            irs[(o, i)] = make_ir(8000 + 200 * o + i, L)
            g.select(o * nin + i, g.filter(irs[(o, i)]))
    xi = interleave([make_noise(8500 + i, 2 * T * B) for i in range(nin)])
    y = run_float(g, xi, T * B)
    assert tensor_ok(g) == 2
    g.close()
    n0 = 2 * T * B - 512
    for o in range(nout):
        want = sum(orc.direct(xi[:, i], irs[(o, i)], n0=n0, count=512) for i in range(nin))
        r = assert_float_parity(y[n0:, o], want, "K = 1024, out %d" % o)
        print("K=1024 out", o, r)
