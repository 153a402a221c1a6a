"""Shared parity helpers for the tests, smoke() and bench.py (TEST INFRASTRUCTURE)."""
import numpy as np

# north_star tolerance for everything that passes through an FFT
SNR_DB_MIN = 110.0
MAX_ABS_REL = 1e-5


def compare_float(got, want):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = got - want
    sig = float((want ** 2).sum())
    noise = float((err ** 2).sum())
    peak = float(np.abs(want).max()) if want.size else 0.0
    snr = float("inf") if noise == 0.0 else (10.0 * np.log10(sig / noise) if sig > 0 else -float("inf"))
    return {"snr_db": snr, "max_abs": float(np.abs(err).max()) if err.size else 0.0, "peak": peak}


def assert_float_parity(got, want, what=""):
    """SNR >= 110 dB and max-abs error <= 1e-5 x peak (BASELINE.json north_star)."""
    r = compare_float(got, want)
    assert r["snr_db"] >= SNR_DB_MIN, "%s: SNR %.1f dB < %.0f dB (%s)" % (what, r["snr_db"], SNR_DB_MIN, r)
    assert r["max_abs"] <= MAX_ABS_REL * max(r["peak"], 1e-30), "%s: max-abs %.3g > 1e-5 x peak %.3g" % (
        what, r["max_abs"], r["peak"])
    return r


def s24_to_float(b):
    """interleaved little-endian 24-bit PCM bytes -> float64 in [-1, 1)."""
    b = np.asarray(b, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    v = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
    v = np.where(v & 0x800000, v - (1 << 24), v)
    return v.astype(np.float64) / float(1 << 23)


def float_to_s24_bytes(x, lib):
    """float32 array -> 24-bit LE bytes through the given CPU library's TransferSamples law."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    out = np.zeros(x.size * 3, dtype=np.uint8)
    lib.transfer(x.view(np.uint8), 4, False, 0, x.size, out, 2, False, 0, x.size, x.size, 1)
    return out
