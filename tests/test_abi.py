"""The C-ABI boundary without a GPU: libbbx.so loads, exports every symbol include/bbx.h declares, the pure
host entry points agree with the oracle, and compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np

import bbcat_dsp_b200 as bbx
import cpulibs as cl
from conftest import ROOT, golden


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bbx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bbx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(bbx.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), "libbbx.so does not export %s" % n
    # and the Python binding binds exactly that set
    assert sorted(bbx.SYMBOLS) == names


def test_version_and_host_helpers(orc):
    assert bbx.lib().bbx_version() == 0x000100
    assert [bbx.GetBytesPerSample(f) for f in range(1, 6)] == [orc.bytes_per_sample(f) for f in range(1, 6)]
    assert [bbx.GetBitsPerSample(f) for f in range(1, 6)] == [orc.bits_per_sample(f) for f in range(1, 6)]
    assert bbx.FractionalSampleAdditionalDelayRequired() == orc.frac_additional() == 14


def test_sanity_checks_match_reference_golden():
    g = golden("formats.npz")
    for case, want in zip(g["sanity_cases"], g["sanity_results"]):
        ok, v = bbx.BlockTransferSanityChecks(*[int(x) for x in case[:6]], allowsinglechannel=bool(case[6]))
        assert int(ok) == want[0]
        if ok:
            assert list(v) == list(want[1:])


def test_interpolator_step_matches_oracle(orc):
    rng = np.random.default_rng(1)
    for _ in range(50):
        a = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1)], dtype=np.float32)
        b = a.copy()
        inc, n = float(np.float32(rng.uniform(0, 0.3))), int(rng.integers(0, 9))
        bbx.InterpolatorStep(a, inc, n)
        orc.interp_step(b, inc, n)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_shard_range_covers_all_channels():
    for n in (1, 2, 7, 32, 64, 128):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                first, count = bbx.shard_range(n, r, world)
                got.extend(range(first, first + count))
            assert got == list(range(n))
    assert bbx.shard_range(128, 1, 8) == (16, 16)


def test_degenerate_transfers_are_noops_even_without_gpu():
    dst = np.full(4, 7.0, dtype=np.float32)
    bbx.TransferSamples(np.ones(4, dtype=np.float32), 4, False, 0, 0, dst, 4, False, 0, 2, 2, 2)  # src_channels == 0
    bbx.MixSamples(np.ones(4, dtype=np.float32), 0, 2, dst, 0, 2, 2, 2, 0.0)  # mul == 0
    assert (dst == 7.0).all()


def test_compute_fails_loudly_without_a_device():
    """No CPU fallback: on a box without a GPU every compute entry point raises with the CUDA reason."""
    if bbx.device_count() > 0:
        return  # on the GPU box the -m gpu tests cover the compute path
    import pytest
    with pytest.raises(bbx.BbxError):
        bbx.Convolver(64, 2, 1)
    with pytest.raises(bbx.BbxError):
        bbx.TransferSamples(np.ones(4, dtype=np.float32), 4, False, 0, 2, np.zeros(4, dtype=np.float32), 4, False, 0, 2, 2, 2)
    with pytest.raises(bbx.BbxError):
        bbx.SoundDelayBuffer()
    assert b"CUDA" in bbx.lib().bbx_last_error()
