#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE'S OWN CODE (oracle/_ref/libbbcref.so, compiled from
/root/reference/src by oracle/Makefile) and, for the convolver (absent from the reference tree), from
float64 direct convolution.  Run in the dev container only; the fixtures travel, the reference does not.

    python tools/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpulibs as cl  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def special_floats():
    v = [0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-9, -1e-9, 3e-10, -3e-10, 0.99999994, -0.99999994, 2.0, -2.0,
         2.0 ** -23, -2.0 ** -23, 1.5 * 2.0 ** -23, -1.5 * 2.0 ** -23, 0.123456789, -0.123456789,
         2.0 ** -31, -2.0 ** -31, 2.0 ** -32, -2.0 ** -32, 1.0 - 2.0 ** -24, -(1.0 + 2.0 ** -23), 1e10, -1e10,
         np.inf, -np.inf, 1e-40, -1e-40, 0.999999999, -0.999999999]
    return np.array(v, dtype=np.float64)


def gen_formats(ref):
    rng = np.random.default_rng(20261018)
    d = {}
    sp64 = special_floats()
    sp32 = sp64.astype(np.float32)
    d["special_f32"] = sp32
    d["special_f64"] = sp64
    for srcname, src, sf in (("f32", sp32, cl.FMT_FLOAT), ("f64", sp64, cl.FMT_DOUBLE)):
        for df in (cl.FMT_16, cl.FMT_24, cl.FMT_32, cl.FMT_FLOAT, cl.FMT_DOUBLE):
            for be in (0, 1):
                dst = np.zeros(src.size * cl.FMT_BYTES[df], dtype=np.uint8)
                ref.transfer(src.view(np.uint8), sf, 0, 0, src.size, dst, df, be, 0, src.size, src.size, 1)
                d["special_%s_to_%s_%s" % (srcname, cl.FMT_NAMES[df], "be" if be else "le")] = dst
    # every table entry on random data, with a channel rectangle
    nfr, sch, dch, nch, sc, dc = 23, 5, 6, 3, 1, 2
    d["rect_geom"] = np.array([nfr, sch, dch, nch, sc, dc], dtype=np.uint32)
    for sf in range(1, 6):
        sb = cl.FMT_BYTES[sf]
        if sf == cl.FMT_FLOAT:
            raw = (rng.standard_normal(nfr * sch) * 0.6).astype("<f4").view(np.uint8)
        elif sf == cl.FMT_DOUBLE:
            raw = (rng.standard_normal(nfr * sch) * 0.6).astype("<f8").view(np.uint8)
        else:
            raw = rng.integers(0, 256, nfr * sch * sb, dtype=np.uint8)
        raw = raw.copy()
        for sbe in (0, 1):
            src = raw.reshape(-1, sb)[:, ::-1].copy().reshape(-1) if sbe else raw
            d["rect_src_%s_%s" % (cl.FMT_NAMES[sf], "be" if sbe else "le")] = src
            for df in range(1, 6):
                for dbe in (0, 1):
                    dst = np.full(nfr * dch * cl.FMT_BYTES[df], 0xA5, dtype=np.uint8)
                    ref.transfer(src, sf, sbe, sc, sch, dst, df, dbe, dc, dch, nch, nfr)
                    d["rect_%s_%s_to_%s_%s" % (cl.FMT_NAMES[sf], "be" if sbe else "le", cl.FMT_NAMES[df],
                                               "be" if dbe else "le")] = dst
    # all 2^16 top patterns of int24 -> float -> int24 (the full 2^24 sweep runs live against _ref in tests)
    v = (np.arange(0, 1 << 24, 251, dtype=np.int64) & 0xFFFFFF).astype(np.uint32)
    b = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).reshape(-1)
    f = np.zeros(v.size, dtype=np.float32)
    ref.transfer(b, cl.FMT_24, 0, 0, v.size, f.view(np.uint8), cl.FMT_FLOAT, 0, 0, v.size, v.size, 1)
    d["s24_sweep_bytes"] = b
    d["s24_sweep_float"] = f
    # sanity-check table
    cases, results = [], []
    for _ in range(400):
        c = [int(rng.integers(0, 6)), int(rng.integers(0, 6)), int(rng.integers(0, 6)), int(rng.integers(0, 6)),
             int(rng.choice([0, 1, 2, 3, 5, 0xFFFFFFFF])), int(rng.integers(0, 4))]
        for allow in (0, 1):
            ok, v = ref.sanity(*c, allow=bool(allow))
            cases.append(c + [allow])
            results.append([int(ok)] + list(v))
    d["sanity_cases"] = np.array(cases, dtype=np.uint32)
    d["sanity_results"] = np.array(results, dtype=np.uint32)
    np.savez_compressed(os.path.join(OUT, "formats.npz"), **d)


def gen_mix(ref):
    rng = np.random.default_rng(31)
    d = {}
    nfr, sch, dch, nch, sc, dc = 19, 3, 4, 2, 1, 1
    d["geom"] = np.array([nfr, sch, dch, nch, sc, dc], dtype=np.uint32)
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        src = rng.standard_normal(nfr * sch).astype(dt)
        dst0 = rng.standard_normal(nfr * dch).astype(dt)
        d["src_" + name] = src
        d["dst0_" + name] = dst0
        for mul in (0.5, 1.0, -0.333, 0.0):
            dst = dst0.copy()
            ref.mix(src, sc, sch, dst, dc, dch, nch, nfr, mul)
            d["mix_%s_%g" % (name, mul)] = dst
    # interpolated mixes: up-ramp, down-ramp, clamp at target, zero/zero no-op
    src = d["src_f32"]
    ramps = np.array([[1.0, 0.0, 0.25], [0.0, 1.0, 0.1], [0.3, 0.0, 0.07], [0.0, 0.0, 0.5], [0.5, 0.5, 0.1]],
                     dtype=np.float32)
    d["ramps"] = ramps
    for i, (target, current, inc) in enumerate(ramps):
        dst = d["dst0_f32"].copy()
        st = np.array([target, current], dtype=np.float32)
        ref.mix_interp(src, sc, sch, dst, dc, dch, nch, nfr, st, float(inc))
        d["ramp_dst_%d" % i] = dst
        d["ramp_state_%d" % i] = st
    np.savez_compressed(os.path.join(OUT, "mix.npz"), **d)


def gen_frac(ref):
    rng = np.random.default_rng(47)
    d = {}
    # impulse scan: every coefficient of the 14 x 128 table is read back exactly once
    length = 64
    buf = np.zeros(length, dtype=np.float32)
    buf[20] = 1.0
    pos = 20.0 + np.arange(0, 15 * 128 + 64) / 128.0
    d["impulse_pos"] = pos
    d["impulse_out"] = ref.frac(buf, 0, 1, length, pos)
    # random interleaved buffers, float and double
    channels, length = 3, 97
    bf = rng.standard_normal(channels * length).astype(np.float32)
    bd = rng.standard_normal(channels * length).astype(np.float64)
    pos = rng.uniform(0, length, 2000)
    pos[:8] = [0.0, 0.5, 13.999, 14.0, length - 1e-9, length - 1.0, 1.0 / 128, 127.0 / 128]
    d["rand_geom"] = np.array([channels, length], dtype=np.uint32)
    d["rand_buf_f32"] = bf
    d["rand_buf_f64"] = bd
    d["rand_pos"] = pos
    for ch in range(channels):
        d["rand_out_f32_ch%d" % ch] = ref.frac(bf, ch, channels, length, pos)
        d["rand_out_f64_ch%d" % ch] = ref.frac(bd, ch, channels, length, pos)
    d["additional"] = np.array([ref.frac_additional()], dtype=np.uint32)
    np.savez_compressed(os.path.join(OUT, "frac.npz"), **d)


def delay_script(seed, n_ops=60):
    """Deterministic random op sequence exercised against SoundDelayBuffer implementations."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(n_ops):
        kind = rng.choice(["write", "inc", "read", "write", "read"])
        if kind == "write":
            ops.append(("write", int(rng.integers(1, 5)), int(rng.integers(0, 4)), int(rng.integers(1, 4)),
                        int(rng.integers(1, 40))))
        elif kind == "inc":
            ops.append(("inc", int(rng.integers(0, 50))))
        else:
            ops.append(("read", int(rng.integers(1, 6)), int(rng.integers(0, 60)), int(rng.integers(0, 4)),
                        int(rng.integers(1, 4)), int(rng.integers(1, 50))))
    return ops


def run_ring_script(lib, seed, chans=3, length=37, fmt=cl.FMT_FLOAT, n_ops=120):
    """Random op sequence against SoundRingBuffer: writes / reads limited by the read position, both increments, a
    resize; every op records its return value and the three position getters."""
    rng = np.random.default_rng(seed)
    d = lib.delay(ring=True)
    d.set_size(chans, length, fmt)
    trace = []

    def pos():
        trace.append(np.array([d.write_position, d.read_position, d.read_available, d.write_available], dtype=np.float64))

    pos()
    for k in range(n_ops):
        kind = rng.choice(["write", "winc", "read", "rinc", "write", "winc", "read"])
        if kind == "write":
            sfmt, ch, nch, nfr = int(rng.integers(1, 5)), int(rng.integers(0, 4)), int(rng.integers(1, 4)), int(rng.integers(1, 50))
            if sfmt >= cl.FMT_FLOAT:
                src = (rng.standard_normal(nfr * chans) * 0.5).astype("<f4" if sfmt == cl.FMT_FLOAT else "<f8").view(np.uint8)
            else:
                src = rng.integers(0, 256, nfr * chans * cl.FMT_BYTES[sfmt], dtype=np.uint8)
            trace.append(np.array([d.write(src.copy(), sfmt, ch, nch, nfr)], dtype=np.float64))
        elif kind == "winc":
            d.increment(int(rng.integers(0, 25)))
        elif kind == "rinc":
            d.increment_read(int(rng.integers(0, 25)))
        else:
            dfmt, delay, ch, nch, nfr = int(rng.integers(1, 6)), int(rng.integers(0, 60)), int(rng.integers(0, 4)), int(rng.integers(1, 4)), int(rng.integers(1, 50))
            dst = np.full(nfr * chans * cl.FMT_BYTES[dfmt], 0x5A, dtype=np.uint8)
            got = d.read(dst, dfmt, delay, ch, nch, nfr)
            trace.append(np.concatenate([[got], dst.astype(np.float64)]))
        pos()
        if k == n_ops // 2:
            d.set_size(chans, length + 11, fmt)  # grow: contents kept, positions stay valid
            pos()
    trace.append(d.raw().astype(np.float64))
    d.close()
    return np.concatenate(trace)


def run_delay_script(lib, seed, chans=3, length=37, fmt=cl.FMT_FLOAT):
    rng = np.random.default_rng(seed + 1000)
    d = lib.delay()
    d.set_size(chans, length, fmt)
    trace = []
    for op in delay_script(seed):
        if op[0] == "write":
            _, sfmt, ch, nch, nfr = op
            n = nfr * min(nch, chans)  # upper bound of samples consumed
            if sfmt >= cl.FMT_FLOAT:
                src = (rng.standard_normal(nfr * chans) * 0.5).astype("<f4" if sfmt == cl.FMT_FLOAT else "<f8").view(np.uint8)
            else:
                src = rng.integers(0, 256, nfr * chans * cl.FMT_BYTES[sfmt], dtype=np.uint8)
            got = d.write(src.copy(), sfmt, ch, nch, nfr)
            trace.append(np.array([got], dtype=np.float64))
            del n
        elif op[0] == "inc":
            d.increment(op[1])
            trace.append(np.array([d.write_position], dtype=np.float64))
        else:
            _, dfmt, delay, ch, nch, nfr = op
            dst = np.full(nfr * chans * cl.FMT_BYTES[dfmt], 0x5A, dtype=np.uint8)
            got = d.read(dst, dfmt, delay, ch, nch, nfr)
            trace.append(np.concatenate([[got], dst.astype(np.float64)]))
    trace.append(d.raw().astype(np.float64))
    d.close()
    return np.concatenate(trace)


def gen_delay(ref):
    d = {}
    for seed in (1, 2, 3):
        for fmt in (cl.FMT_FLOAT, cl.FMT_16, cl.FMT_DOUBLE):
            d["trace_seed%d_fmt%d" % (seed, fmt)] = run_delay_script(ref, seed, fmt=fmt)
    np.savez_compressed(os.path.join(OUT, "delay.npz"), **d)


def gen_ring(ref):
    d = {}
    for seed in (21, 22, 23):
        for fmt in (cl.FMT_FLOAT, cl.FMT_16):
            d["trace_seed%d_fmt%d" % (seed, fmt)] = run_ring_script(ref, seed, fmt=fmt)
    np.savez_compressed(os.path.join(OUT, "ring.npz"), **d)


def conv_case(seed, L, B, nblk, nch=1):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal((nch, L)) * np.exp(-6.9 * np.arange(L) / L)
    h /= np.sqrt((h ** 2).sum(axis=1, keepdims=True))
    h = h.astype(np.float32)
    x = rng.uniform(-1, 1, (nch, nblk * B)).astype(np.float32)
    return h, x


def gen_conv():
    """float64 truth for the absent BlockConvolver: y = h * x by direct summation (numpy float64)."""
    d = {}
    cases = [(101, 8192, 1024, 12), (102, 512, 256, 10), (103, 4096, 512, 12), (104, 200, 64, 9), (105, 1000, 128, 11)]
    d["cases"] = np.array(cases, dtype=np.uint32)
    for seed, L, B, nblk in cases:
        h, x = conv_case(seed, L, B, nblk)
        y = np.convolve(x[0].astype(np.float64), h[0].astype(np.float64))[: nblk * B]
        d["y_%d" % seed] = y
    np.savez_compressed(os.path.join(OUT, "conv_truth.npz"), **d)


# ---- BiQuadCoeffs / BiQuad (src/BiQuad.cpp), SURVEY 8f.4 ----
BIQUAD_DESIGNS = [(t, f, fs, g, bw) for t in range(10)
                  for (f, fs, g, bw) in ((1000.0, 48000.0, 6.0, 1.0), (120.5, 44100.0, -9.5, 0.7), (8000.0, 48000.0, 3.0, 2.5))]


def biquad_script(lib, nch=5, nsrc=7, ndst=6, seed=4242):
    """One scenario run on any implementation: jumps, ramps given in samples (SetCoeffs) and in seconds (CalcCoeffs),
    a ramp that ends inside a call, a retarget in the middle of a ramp, partial channel counts, in-place processing,
    Reset.  Returns every output block and the state after every call."""
    rng = np.random.default_rng(seed)
    bq = lib.biquad(nch)
    outs = []

    def run(nframes, nchannels=nch, s=nsrc, d=ndst, inplace=False):
        x = rng.uniform(-1, 1, nframes * s).astype(np.float32)
        if inplace:
            y = x.copy()
            bq.process(y, y, nchannels, s, s, nframes)
        else:
            y = np.full(nframes * d, 7.0, dtype=np.float32)
            bq.process(x, y, nchannels, s, d, nframes)
        w, cur, md = bq.state()
        outs.extend([y, w.copy(), cur.copy(), md.copy()])

    run(33)                                                    # default-constructed coeffs: flat
    bq.calc(7, 1000.0, 48000.0, 6.0, 1.0, 0.0)                 # PEQ, immediate
    run(100)
    bq.calc(3, 3000.0, 48000.0, 0.0, 1.0, 0.004)               # LPF12, 4 ms = 192-sample ramp
    run(64)
    run(64, nchannels=3)                                       # two filters keep their state, ramp continues
    run(200)                                                   # ramp ends inside this call
    bq.set_coeffs(lib.biquad_coeffs(8, 250.0, 48000.0, -4.0, 1.0), 37.5)   # LSH, 37.5-sample ramp
    run(20)
    bq.calc(6, 5000.0, 48000.0, 0.0, 0.3, 0.001)               # retarget in the middle of the ramp
    run(90, inplace=True)
    bq.reset()
    run(17, nchannels=99)                                      # clamped to the channels that exist
    bq.close()
    return outs


def gen_biquad(ref):
    d = {"designs": np.array(BIQUAD_DESIGNS, dtype=np.float64)}
    d["coeffs"] = np.stack([ref.biquad_coeffs(int(t), f, fs, g, bw) for (t, f, fs, g, bw) in BIQUAD_DESIGNS])
    for i, a in enumerate(biquad_script(ref)):
        d["script_%03d" % i] = a
    np.savez_compressed(os.path.join(OUT, "biquad.npz"), **d)


# ---- BiQuadFilterBank (src/BiQuad.h:247-353, src/BiQuad.cpp:498-662) ----
def fbank_script(lib, nch=5, nfilters=3, nsrc=7, ndst=6, seed=9191):
    """One scenario on any implementation: filters designed and set explicitly, ramps on some filters only, a ramp that ends
    inside a call, partial channel counts, in-place processing, AddFilter / SetFilters / SetChannels between calls, Reset.
    Returns every output block and, after every call, the state of every filter."""
    rng = np.random.default_rng(seed)
    fb = lib.fbank(nch, nfilters)
    outs = []
    size = [nch, nfilters]

    def run(nframes, nchannels=None, s=nsrc, d=ndst, inplace=False):
        nchannels = size[0] if nchannels is None else nchannels
        x = rng.uniform(-1, 1, nframes * s).astype(np.float32)
        if inplace:
            y = x.copy()
            fb.process(y, y, nchannels, s, s, nframes)
        else:
            y = np.full(nframes * d, 7.0, dtype=np.float32)
            fb.process(x, y, nchannels, s, d, nframes)
        outs.append(y)
        for f in range(size[1]):
            w, cur, md = fb.state(f)
            outs.extend([w.copy(), cur.copy(), md.copy()])

    run(21)                                                       # default filters: flat
    for f in range(nfilters):
        fb.calc(f, (7, 8, 9, 3, 4, 5, 6)[f % 7], 300.0 * (f + 1), 48000.0, 3.0 - f, 1.0, 0.0)
    run(100)
    fb.calc(1 % nfilters, 3, 2500.0, 48000.0, 0.0, 1.0, 0.003)  # one filter ramps for 144 samples, the others stay
    run(50)
    run(61, nchannels=3)                                          # two channels keep their state, the ramp goes on
    fb.set_coeffs(0, lib.biquad_coeffs(8, 200.0, 48000.0, -5.0, 1.0), 29.5)   # a second ramp while the first still runs
    run(70)                                                       # both end inside this call
    run(40, inplace=True)
    fb.add_filter(lib.biquad_coeffs(7, 4000.0, 48000.0, 4.0, 0.7))
    size[1] += 1
    run(33)
    fb.set_channels(nch + 2)                                      # new channels start silent, old ones keep their state
    size[0] = nch + 2
    run(25, s=nch + 3, d=nch + 2)
    fb.set_filters(2)                                             # filters leave from the end
    size[1] = 2
    run(19, s=nch + 3, d=nch + 2)
    fb.set_channels(2)
    size[0] = 2
    run(16, nchannels=99)                                         # clamped to the channels that exist
    fb.reset()
    run(9)
    fb.close()
    return outs


def gen_fbank(ref):
    d = {}
    for i, a in enumerate(fbank_script(ref)):
        d["script_%03d" % i] = a
    for i, a in enumerate(fbank_script(ref, nch=70, nfilters=19, nsrc=72, ndst=71, seed=77)):   # more than one pass of 16
        d["long_%03d" % i] = a
    np.savez_compressed(os.path.join(OUT, "fbank.npz"), **d)


# ---- AllPassFilterChain<float> (src/AllPassFilter.h), SURVEY 8f.4 ----
def allpass_script(lib, nch=5, delays=(7, 1, 23, 4), coeffs=(0.5, -0.7, 0.3, 0.9), nsrc=8, ndst=6, seed=777, offs=((1, 1), (2, 3))):
    """Chain processed in several calls (ring wrap inside a call, positions carried over), a channel offset, a geometry
    that fits only some of the channels (the others are skipped with Advance), in-place.  Returns outputs and states."""
    rng = np.random.default_rng(seed)
    ap = lib.allpass(nch, delays, coeffs)
    outs = []

    def run(nframes, sc=0, dc=0, s=nsrc, d=ndst, inplace=False):
        x = rng.uniform(-1, 1, nframes * s).astype(np.float32)
        if inplace:
            y = x.copy()
            ap.process(y, y, sc, s, sc, s, nframes)
        else:
            y = np.full(nframes * d, 3.0, dtype=np.float32)
            ap.process(x, y, sc, s, dc, d, nframes)
        outs.append(y)
        for f in range(len(delays)):
            ring, pos = ap.state(f)
            outs.extend([ring.copy(), np.array([pos], dtype=np.uint32)])

    run(40)
    run(3, sc=offs[0][0], dc=offs[0][1])    # default geometry: 5 channels fit from offset 1
    run(50, sc=offs[1][0], dc=offs[1][1])   # default geometry: only 3 channels fit the destination, two are skipped
    run(31, inplace=True)
    ap.close()
    return outs


ALLPASS_NARROW = dict(nch=6, delays=(3, 5, 2), coeffs=(0.5, -0.6, 0.8), nsrc=4, ndst=7, seed=779, offs=((0, 0), (1, 1)))


def gen_allpass(ref):
    d = {}
    for i, a in enumerate(allpass_script(ref)):
        d["script_%03d" % i] = a
    # single channel (the reference's unclamped branch): offsets must stay inside the frames
    for i, a in enumerate(allpass_script(ref, nch=1, delays=(5,), coeffs=(0.6,), nsrc=3, ndst=2, seed=778, offs=((1, 1), (2, 0)))):
        d["mono_%03d" % i] = a
    # source narrower than the destination: the first section fits 4 channels, the following ones (dst geometry) all 6
    for i, a in enumerate(allpass_script(ref, **ALLPASS_NARROW)):
        d["narrow_%03d" % i] = a
    np.savez_compressed(os.path.join(OUT, "allpass.npz"), **d)


# ---- BiQuadCascade (src/BiQuad.h:373-792), SURVEY 8f.4 ----
def cascade_coeffs(rng, nf):
    """stable sections: poles inside the unit circle (|a2| < 1, |a1| < 1 + a2), arbitrary zeros; g is never applied"""
    v = [0.5]
    for _ in range(nf):
        a2 = rng.uniform(-0.6, 0.9)
        a1 = rng.uniform(-1, 1) * (1 + a2) * 0.95
        v += [rng.uniform(-1.5, 1.5), rng.uniform(-1, 1), a1, a2]
    return np.array(v, dtype=np.float32)


def cascade_script(lib, nch=5, nf=8, vectorise=True, seed=4242):
    """A bank processed in several calls (registers carried over): shared coefficients, one channel re-programmed (which
    resets only that channel), planar and interleaved layouts, Reset.  Returns outputs and register states."""
    rng = np.random.default_rng(seed)
    cs = lib.cascade(nch, nf, vectorise, True)
    outs = []

    def snap():
        for j in (0, nch - 1):
            x, y, w0, w1, last, info = cs.state(j)
            outs.extend([x[:nf].copy(), y[:nf].copy(), w0[:nf].copy(), w1[:nf].copy(), last.copy(), np.array([info], dtype=np.uint32)])

    def run(nframes, interleaved=True):
        x = rng.uniform(-1, 1, nframes * nch).astype(np.float32)
        y = np.full(nframes * nch, 3.0, dtype=np.float32)
        cs.process(x, y, nframes, interleaved)
        outs.append(y)
        snap()

    run(9)                                             # default pass-through (delayed by nf - 1 samples when vectorised)
    assert cs.set_coefficients(cascade_coeffs(rng, nf))
    assert not cs.set_coefficients(np.zeros(4 * nf, dtype=np.float32))   # wrong length: rejected, nothing changes
    run(50)
    run(7, interleaved=False)
    assert cs.set_coefficients(cascade_coeffs(rng, nf), channel=nch - 1)  # resets the registers of that channel only
    run(33)
    cs.reset()
    run(1)
    run(18)
    cs.close()
    return outs


def gen_cascade(ref):
    d = {}
    for name, kw in CASCADE_CASES.items():
        for i, a in enumerate(cascade_script(ref, **kw)):
            d["%s_%03d" % (name, i)] = a
    np.savez_compressed(os.path.join(OUT, "cascade.npz"), **d)


CASCADE_CASES = {
    "vec8": dict(nch=5, nf=8, vectorise=True, seed=4242),
    "vec12": dict(nch=3, nf=12, vectorise=True, seed=4243),
    "plain5": dict(nch=4, nf=5, vectorise=False, seed=4244),
    "novec6": dict(nch=2, nf=6, vectorise=True, seed=4245),   # 6 % 4 != 0: the reference switches vectorise off
    "one": dict(nch=1, nf=1, vectorise=False, seed=4246),
}


def gen_dither(ref):
    """the reference's TransferSamples driven with the stateful test Ditherer (tests/cpp/test_ditherer.h), every converter
    with a dither call site, both byte orders, three geometries (tests/test_dither.py)"""
    import test_dither as td
    d = {}
    for s, dd, src_be, dst_be, gi, geom in td.all_cases():
        out, calls = td.run_case(ref.transfer_ditherer, s, dd, src_be, dst_be, geom, 1)
        assert calls == geom[4] * geom[5]
        d["hook_%d_%d_%d_%d_%d" % (s, dd, src_be, dst_be, gi)] = out
    np.savez_compressed(os.path.join(OUT, "dither.npz"), **d)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = cl.reference()
    if ref is None:
        sys.exit("oracle/_ref/libbbcref.so missing: run `make -C oracle` in the dev container first")
    gen_formats(ref)
    gen_mix(ref)
    gen_frac(ref)
    gen_delay(ref)
    gen_ring(ref)
    gen_biquad(ref)
    gen_fbank(ref)
    gen_allpass(ref)
    gen_cascade(ref)
    gen_dither(ref)
    gen_conv()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
