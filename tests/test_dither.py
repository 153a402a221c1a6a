"""a7 Ditherer hook (src/SoundFormatConversions.h:39-54) -- SURVEY.md 8(a) row a7.

Three things are pinned:
  * WHERE the reference calls ditherer->Dither and WITH WHAT (frame loop counter, bit count, call order): the oracle's hook
    form against the reference's own TransferSamples driven with the stateful test subclass of tests/cpp/test_ditherer.h
    (live when oracle/_ref is present, golden vectors from it otherwise), for every converter and both frame directions;
  * the C++ host shim honours an arbitrary subclass exactly like that (GPU, through libbbx): a no-op subclass gives the
    NULL result byte for byte, the test subclass gives the reference's bytes;
  * Dither_TPDF, the device option (libbbx's own law -- the reference names the mode and ships no implementation, so
    parity is pinned to oracle/formats.c only): GPU bytes == oracle bytes, and the noise has the stated statistics.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import cpulibs as cl
from conftest import ROOT, golden

FMTS = [cl.FMT_16, cl.FMT_24, cl.FMT_32, cl.FMT_FLOAT, cl.FMT_DOUBLE]
EXPECT_BITS = {(2, 1): 16, (3, 1): 16, (4, 1): 16, (5, 1): 16, (3, 2): 8, (4, 2): 8, (5, 2): 8, (5, 3): 0}
# (src_channel, src_channels, dst_channel, dst_channels, nchannels, nframes): contiguous (collapses to one frame), strided
# forwards, and a destination frame longer than the source frame (the reference then runs the frames backwards)
GEOMS = [(0, 3, 0, 3, 3, 7), (1, 4, 0, 3, 2, 9), (0, 2, 3, 11, 2, 6)]


def source_bytes(fmt, n, seed):
    rng = np.random.default_rng(seed)
    if fmt == cl.FMT_FLOAT:
        return (rng.uniform(-1.1, 1.1, n).astype(np.float32)).view(np.uint8).copy()
    if fmt == cl.FMT_DOUBLE:
        return rng.uniform(-1.1, 1.1, n).view(np.uint8).copy()
    return rng.integers(0, 256, n * cl.FMT_BYTES[fmt], dtype=np.uint8)


def run_case(fn, src_fmt, dst_fmt, src_be, dst_be, geom, mode, seed=11):
    sc, scs, dc, dcs, nch, nfr = geom
    src = source_bytes(src_fmt, scs * nfr, seed)
    if src_be and src_fmt >= cl.FMT_FLOAT:  # byte-swapped floating point source
        w = cl.FMT_BYTES[src_fmt]
        src = src.reshape(-1, w)[:, ::-1].reshape(-1).copy()
    dst = np.full(dcs * nfr * cl.FMT_BYTES[dst_fmt], 0xA5, dtype=np.uint8)
    calls = fn(src, src_fmt, src_be, sc, scs, dst, dst_fmt, dst_be, dc, dcs, nch, nfr, mode)
    return dst, calls


def all_cases():
    for (s, d) in sorted(EXPECT_BITS):
        for src_be in (False, True):
            for dst_be in (False, True):
                for gi, geom in enumerate(GEOMS):
                    yield s, d, src_be, dst_be, gi, geom


def test_dither_call_site_table(orc, bbx_lib_cpu):
    for s in FMTS:
        for d in FMTS:
            want = EXPECT_BITS.get((s, d), -1)
            assert orc.dither_bits(s, d) == want
            assert bbx_lib_cpu.bbx_dither_bits(s, d) == want


@pytest.fixture(scope="module")
def bbx_lib_cpu():
    """the product library loads without a GPU (host-only entry points such as bbx_dither_bits work)"""
    import bbcat_dsp_b200 as b
    return b.lib()


def test_oracle_hook_matches_reference_live(orc, ref):
    for s, d, src_be, dst_be, gi, geom in all_cases():
        for mode in (0, 1):
            a, ca = run_case(orc.transfer_ditherer, s, d, src_be, dst_be, geom, mode)
            b, cb = run_case(ref.transfer_ditherer, s, d, src_be, dst_be, geom, mode)
            assert ca == cb and np.array_equal(a, b), (s, d, src_be, dst_be, gi, mode)
            if mode == 1:
                assert ca == geom[4] * geom[5]  # one Dither() call per converted sample
    # converters without a call site never touch the ditherer
    for s, d in ((1, 2), (1, 4), (4, 5), (4, 3), (2, 2), (4, 4)):
        b, cb = run_case(ref.transfer_ditherer, s, d, False, False, GEOMS[1], 1)
        a, ca = run_case(orc.transfer_ditherer, s, d, False, False, GEOMS[1], 1)
        assert ca == cb == 0 and np.array_equal(a, b)


def test_oracle_hook_matches_golden(orc):
    g = golden("dither.npz")
    n = 0
    for s, d, src_be, dst_be, gi, geom in all_cases():
        a, _ = run_case(orc.transfer_ditherer, s, d, src_be, dst_be, geom, 1)
        assert np.array_equal(a, g["hook_%d_%d_%d_%d_%d" % (s, d, src_be, dst_be, gi)])
        n += 1
    assert n == 8 * 4 * 3


def test_noop_ditherer_equals_plain_transfer(orc):
    for s, d, src_be, dst_be, gi, geom in all_cases():
        a, _ = run_case(orc.transfer_ditherer, s, d, src_be, dst_be, geom, 0)
        b, _ = run_case(lambda *x: orc.transfer(*x[:-1]), s, d, src_be, dst_be, geom, 0)
        assert np.array_equal(a, b)


# ---- GPU: the C++ shim over libbbx, and the device TPDF option -------------------------------------------------------
@pytest.fixture(scope="module")
def shim(tmp_path_factory, bbx):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    so = str(tmp_path_factory.mktemp("shim") / "libshim_dither.so")
    libdir = os.path.join(ROOT, "bbcat-dsp_b200")
    subprocess.check_call([cxx, "-std=c++11", "-Wall", "-Werror", "-fPIC", "-shared", "-I" + os.path.join(libdir, "host"),
                           "-I" + os.path.join(ROOT, "tests", "cpp"), os.path.join(ROOT, "tests", "cpp", "shim_dither.cpp"),
                           "-o", so, "-L" + libdir, "-lbbx", "-Wl,-rpath," + libdir])
    L = C.CDLL(so)
    L.shim_transfer_samples_ditherer.restype = C.c_uint
    L.shim_transfer_samples_ditherer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_int, C.c_int,
                                                 C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_int]
    L.shim_tpdf_first_seed.restype = C.c_ulonglong

    def call(src, st, sbe, sc, scs, dst, dt, dbe, dc, dcs, nch, nfr, mode):
        return L.shim_transfer_samples_ditherer(src.ctypes.data_as(C.c_void_p), st, int(sbe), sc, scs, dst.ctypes.data_as(C.c_void_p),
                                                dt, int(dbe), dc, dcs, nch, nfr, mode)
    call.first_seed = L.shim_tpdf_first_seed()
    return call


@pytest.mark.gpu
def test_shim_honours_ditherer_subclasses(shim, orc, bbx):
    """no-op subclass == NULL result byte for byte; the stateful test subclass == the reference's bytes (oracle hook form,
    pinned to the reference above) with the same number of calls; TPDFDitherer == the device option with its first seed."""
    for s, d, src_be, dst_be, gi, geom in all_cases():
        plain, _ = run_case(lambda *x: orc.transfer(*x[:-1]), s, d, src_be, dst_be, geom, 0)
        a, _ = run_case(shim, s, d, src_be, dst_be, geom, 0)
        assert np.array_equal(a, plain), ("no-op", s, d, src_be, dst_be, gi)
        a, ca = run_case(shim, s, d, src_be, dst_be, geom, 1)
        b, cb = run_case(orc.transfer_ditherer, s, d, src_be, dst_be, geom, 1)
        assert ca == cb and np.array_equal(a, b), ("test ditherer", s, d, src_be, dst_be, gi)
        a, _ = run_case(shim, s, d, src_be, dst_be, geom, 2)
        b, _ = run_case(lambda *x: orc.transfer_tpdf(*x[:-1], shim.first_seed), s, d, src_be, dst_be, geom, 2)
        assert np.array_equal(a, b), ("tpdf", s, d, src_be, dst_be, gi)


@pytest.mark.gpu
def test_shim_matches_reference_live(shim, ref):
    for s, d, src_be, dst_be, gi, geom in all_cases():
        a, ca = run_case(shim, s, d, src_be, dst_be, geom, 1)
        b, cb = run_case(ref.transfer_ditherer, s, d, src_be, dst_be, geom, 1)
        assert ca == cb and np.array_equal(a, b), (s, d, src_be, dst_be, gi)


@pytest.mark.gpu
def test_tpdf_device_equals_oracle(bbx, orc):
    def gpu(src, st, sbe, sc, scs, dst, dt, dbe, dc, dcs, nch, nfr, seed):
        bbx.TransferSamples(src, st, sbe, sc, scs, dst, dt, dbe, dc, dcs, nch, nfr, dither=bbx.DITHER_TPDF, seed=seed)
    for s in FMTS:
        for d in FMTS:
            for geom in GEOMS + [(0, 32, 0, 32, 32, 4096)]:
                for seed in (0, 0xDEADBEEFCAFE):
                    a, _ = run_case(gpu, s, d, False, s == 3, geom, seed)
                    b, _ = run_case(orc.transfer_tpdf, s, d, False, s == 3, geom, seed)
                    assert np.array_equal(a, b), (s, d, geom, seed)


@pytest.mark.parametrize("impl_name", ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def test_tpdf_statistics(impl_name, orc, request):
    """float -> 16-bit with TPDF: the error against the unquantised value is zero-mean, stays inside +-1.5 LSB, and its
    variance is that of TPDF + uniform quantisation (1/6 + 1/12 LSB^2); without dither the truncating converter is biased
    by half an LSB."""
    n = 1 << 16
    x = np.random.default_rng(5).uniform(-0.5, 0.5, n).astype(np.float32)
    dst = np.zeros(n, dtype=np.int16)
    if impl_name == "gpu":
        bbx = request.getfixturevalue("bbx")
        bbx.TransferSamples(x.view(np.uint8), cl.FMT_FLOAT, False, 0, 1, dst.view(np.uint8), cl.FMT_16, False, 0, 1, 1, n,
                            dither=bbx.DITHER_TPDF, seed=3)
    else:
        orc.transfer_tpdf(x.view(np.uint8), cl.FMT_FLOAT, False, 0, 1, dst.view(np.uint8), cl.FMT_16, False, 0, 1, 1, n, 3)
    err = dst.astype(np.float64) - x.astype(np.float64) * 32768.0  # in LSB(16)
    assert abs(err.mean()) < 0.02 and np.abs(err).max() <= 1.5
    assert abs(err.var() - 0.25) < 0.02
    plain = np.zeros(n, dtype=np.int16)
    orc.transfer(x.view(np.uint8), cl.FMT_FLOAT, False, 0, 1, plain.view(np.uint8), cl.FMT_16, False, 0, 1, 1, n)
    # two's-complement truncation of the 32-bit word: floor, i.e. -0.5 LSB on average
    assert abs((plain.astype(np.float64) - x.astype(np.float64) * 32768.0).mean() + 0.5) < 0.02
