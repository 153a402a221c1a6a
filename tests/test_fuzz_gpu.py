"""Randomised engine scenarios, CUDA product against the CPU oracle (same seeded script on both): modes (per-channel, routed
mixdown, MIMO), block sizes, filter lengths, ragged call sizes, integer / fractional delays, hard and crossfaded switches
of filter and delay, zero gains, every PCM format with either endianness, extra PCM channels, pinned (latency path) and
pageable host buffers.  Float outputs: the path's tolerance (SNR >= 110 dB, max-abs <= 1e-5 x peak); integer outputs: the
same plus one LSB of the format."""
import os

import numpy as np
import pytest

import cpulibs as cl
from convkit import GpuDriver, OracleDriver, make_ir, make_noise
from parity import compare_float

pytestmark = pytest.mark.gpu

FMTS = (cl.FMT_16, cl.FMT_24, cl.FMT_32, cl.FMT_FLOAT, cl.FMT_DOUBLE)
LSB = {cl.FMT_16: 2.0 ** -15, cl.FMT_24: 2.0 ** -23, cl.FMT_32: 2.0 ** -31}


def _scenario(seed):
    rng = np.random.default_rng(seed)
    mode = [cl.MODE_PER_CHANNEL, cl.MODE_ROUTED, cl.MODE_MIMO][int(rng.integers(0, 3))]
    B = int(rng.choice([64, 128, 256, 512]))
    tmax = int(rng.integers(1, 21))
    sc = dict(seed=seed, mode=mode, B=B, tmax=tmax, fractional=bool(rng.integers(0, 2)), max_delay=int(rng.integers(0, 120)))
    if mode == cl.MODE_PER_CHANNEL:
        sc["nin"] = sc["nout"] = int(rng.integers(1, 41))
        sc["npaths"] = sc["nin"]
    elif mode == cl.MODE_ROUTED:
        sc["nin"], sc["nout"] = int(rng.integers(1, 20)), int(rng.integers(1, 6))
        sc["npaths"] = int(rng.integers(1, 60))
    else:
        sc["nin"], sc["nout"] = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        sc["npaths"] = sc["nin"] * sc["nout"]
        sc["fractional"], sc["max_delay"] = False, 0
    sc["pmax"] = int(rng.integers(1, 7)) if rng.integers(0, 4) else int(rng.integers(32, 45))  # sometimes long: time-batched MAC
    sc["infmt"], sc["outfmt"] = int(rng.choice(FMTS)), int(rng.choice(FMTS))
    sc["in_be"], sc["out_be"] = bool(rng.integers(0, 4) == 0), bool(rng.integers(0, 4) == 0)
    sc["in_extra"], sc["out_extra"] = int(rng.integers(0, 3)), int(rng.integers(0, 3))
    sc["pinned"] = bool(rng.integers(0, 2))
    sc["calls"] = [int(rng.integers(1, tmax + 1)) for _ in range(int(rng.integers(3, 8)))]
    return sc


def _run(drv, sc, orc, bbx=None):
    """the scenario's op script on one driver; returns the concatenated output PCM bytes"""
    rng = np.random.default_rng(sc["seed"] + 10_000)
    B, nin, nout, npaths, mode = sc["B"], sc["nin"], sc["nout"], sc["npaths"], sc["mode"]
    nbank = 3
    irs = [[make_ir(int(rng.integers(1, 1 << 30)), int(rng.integers(1, sc["pmax"] * B + 1))) * 0.3 for _ in range(nbank)]
           for _ in range(npaths)]
    filt = [[drv.filter(h) for h in row] for row in irs]
    inputs = [int(rng.integers(0, nin)) for _ in range(npaths)]
    outputs = [int(rng.integers(0, nout)) for _ in range(npaths)]
    if mode == cl.MODE_ROUTED:
        for p in range(npaths):
            drv.route(p, inputs[p], outputs[p], 0.0 if rng.integers(0, 8) == 0 else float(rng.uniform(0.1, 0.6)))

    def delay():
        if sc["max_delay"] == 0:
            return 0.0
        d = rng.uniform(0, sc["max_delay"])
        return float(d) if sc["fractional"] else float(int(d))

    for p in range(npaths):
        if rng.integers(0, 10):  # some paths stay without a filter (silent)
            drv.select(p, filt[p][0], delay=delay())
    in_ch, out_ch = nin + sc["in_extra"], nout + sc["out_extra"]
    ibps, obps = cl.FMT_BYTES[sc["infmt"]], cl.FMT_BYTES[sc["outfmt"]]
    outs = []
    for k, nb in enumerate(sc["calls"]):
        if k and rng.integers(0, 2):
            for p in rng.choice(npaths, size=min(npaths, int(rng.integers(1, 4))), replace=False):
                xf = bool(rng.integers(0, 2))
                drv.select(int(p), filt[int(p)][int(rng.integers(0, nbank))], delay=delay(), crossfade=xf)
        x = (np.stack([make_noise(int(rng.integers(1, 1 << 30)), nb * B) for _ in range(in_ch)], axis=1) * 0.2).astype(np.float32)
        pcm = np.zeros(x.size * ibps, dtype=np.uint8)
        orc.transfer(np.ascontiguousarray(x).view(np.uint8).reshape(-1), cl.FMT_FLOAT, 0, 0, in_ch, pcm, sc["infmt"], sc["in_be"], 0, in_ch,
                     in_ch, nb * B)
        nout_bytes = nb * B * out_ch * obps
        if drv.name == "gpu" and sc["pinned"]:
            pin, pout = bbx.PinnedBuffer(pcm.size), bbx.PinnedBuffer(nout_bytes)
            pin.array[:] = pcm
            pout.array[:] = 0x11
            from bbcat_dsp_b200 import lib, vp, _check
            _check(lib().bbx_process(drv.eng.h, vp(pin.ptr), sc["infmt"], int(sc["in_be"]), in_ch, vp(pout.ptr), sc["outfmt"],
                                     int(sc["out_be"]), out_ch, nb * B))
            y = pout.array.copy()
        else:
            y = np.full(nout_bytes, 0x11, dtype=np.uint8)
            if drv.name == "gpu":
                drv.eng.Convolve(pcm, sc["infmt"], in_ch, sc["outfmt"], out_ch, nb * B, sc["in_be"], sc["out_be"], out=y)
            else:
                drv.cv.process(pcm, sc["infmt"], in_ch, sc["outfmt"], out_ch, nb * B, sc["in_be"], sc["out_be"], out=y)
        outs.append(np.array(y, copy=True).reshape(nb * B, out_ch, obps))
    return np.concatenate(outs)


@pytest.mark.parametrize("seed", range(int(os.environ.get("BBX_FUZZ_SEEDS", "40"))))  # BBX_FUZZ_SEEDS=N widens the hunt
def test_random_scenario_vs_oracle(bbx, orc, seed):
    sc = _scenario(seed)
    kw = dict(n_outputs=sc["nout"], n_paths=sc["npaths"], mode=sc["mode"], max_blocks=sc["tmax"], max_delay=sc["max_delay"],
              fractional_delay=sc["fractional"])
    g = GpuDriver(bbx, sc["B"], sc["pmax"], sc["nin"], **kw)
    o = OracleDriver(sc["B"], sc["pmax"], sc["nin"], **kw)
    assert g.ring_length == o.ring_length
    yg = _run(g, sc, orc, bbx)
    yo = _run(o, sc, orc)
    g.close()
    nout, obps = sc["nout"], cl.FMT_BYTES[sc["outfmt"]]
    assert (yg[:, nout:] == 0x11).all() and (yo[:, nout:] == 0x11).all(), "channels beyond n_outputs were touched: %s" % sc
    frames = yg.shape[0]
    fa, fb = np.zeros(frames * nout, dtype=np.float64), np.zeros(frames * nout, dtype=np.float64)
    for arr, dst in ((yg, fa), (yo, fb)):
        src = np.ascontiguousarray(arr[:, :nout]).reshape(-1)
        orc.transfer(src, sc["outfmt"], sc["out_be"], 0, nout, dst.view(np.uint8), cl.FMT_DOUBLE, 0, 0, nout, nout, frames)
    r = compare_float(fa, fb)
    assert r["peak"] < 0.999, "scenario clips: %s" % sc
    lsb = LSB.get(sc["outfmt"], 0.0)
    if r["peak"] > 0:
        # integer outputs: the float tolerance plus one LSB where the converter inputs straddle a step;
        # the SNR bound only applies when the quantisation floor is far below it
        assert r["max_abs"] <= lsb + 1e-5 * r["peak"], (r, sc)
        if lsb == 0.0 or lsb < 1e-6 * r["peak"]:
            assert r["snr_db"] >= 110.0, (r, sc)
    else:
        assert r["max_abs"] == 0.0, (r, sc)
