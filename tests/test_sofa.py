"""SOFA impulse-response sets (README:77-78 lists src/SOFA.{h,cpp}, absent from the tree).

The container parser of bbx_sofa_* is pinned against an independent netCDF implementation: scipy.io.netcdf_file writes
the files (CDF-1 and CDF-2, fixed and record dimensions) and reads them back next to libbbx.  Parity against BBC's SOFA
class is unpinned (no source, no libnetcdf here).  The GPU test builds a filter bank from a file and renders through it.
"""

import numpy as np
import pytest
from scipy.io import netcdf_file

import bbcat_dsp_b200 as bbx


def write_sofa(path, ir, delay, rate, src, version=1, record_m=False, fire=False, src_type="spherical", extra_vars=True,
               ir_dtype="d"):
    """ir: [M][R][N] (or [M][R][E][N] with fire), delay: [1 or M][R]([E]), rate: scalar, src: [M][3]."""
    M, R = ir.shape[0], ir.shape[1]
    N = ir.shape[-1]
    f = netcdf_file(path, "w", version=version)
    f.Conventions = "SOFA"
    f.Version = "1.0"
    f.SOFAConventions = "SimpleFreeFieldHRIR"
    f.DataType = "FIRE" if fire else "FIR"
    f.RoomType = "free field"
    f.Title = "libbbx test set"
    if record_m:
        f.createDimension("M", None)  # the record dimension has to come first in scipy's writer
    f.createDimension("I", 1)
    f.createDimension("C", 3)
    f.createDimension("R", R)
    f.createDimension("E", ir.shape[2] if fire else 1)
    f.createDimension("N", N)
    if not record_m:
        f.createDimension("M", M)
    dims = ("M", "R", "E", "N") if fire else ("M", "R", "N")
    v = f.createVariable("Data.IR", ir_dtype, dims)
    v[:] = ir
    v = f.createVariable("Data.SamplingRate", "d", ("I",))
    v[:] = [rate]
    v.Units = "hertz"
    ddims = (("M" if delay.shape[0] == M and M > 1 else "I"), "R") + (("E",) if delay.ndim == 3 else ())
    v = f.createVariable("Data.Delay", "d", ddims)
    v[:] = delay
    v = f.createVariable("SourcePosition", "d", ("M", "C"))
    v[:] = src
    v.Type = src_type
    v.Units = "degree, degree, metre" if src_type == "spherical" else "metre"
    if extra_vars:
        v = f.createVariable("ListenerPosition", "d", ("I", "C"))
        v[:] = [[0.0, 0.0, 0.0]]
        v.Type = "cartesian"
        v.Units = "metre"
        v = f.createVariable("ReceiverPosition", "d", ("R", "C", "I"))
        rp = np.zeros((R, 3, 1))
        rp[:, 1, 0] = np.linspace(0.09, -0.09, R)
        v[:] = rp
        v.Type = "cartesian"
        v = f.createVariable("EmitterPosition", "f", ("E", "C", "I"))
        v[:] = np.zeros((ir.shape[2] if fire else 1, 3, 1), dtype=np.float32)
        v.Type = "cartesian"
        v = f.createVariable("MeasurementIndex", "i", ("M",))  # an int variable the reader must skip over
        v[:] = np.arange(M)
    f.close()


def make_set(rng, M=7, R=2, N=40, E=None):
    shape = (M, R, N) if E is None else (M, R, E, N)
    ir = rng.standard_normal(shape) * np.exp(-np.arange(N) / 9.0)
    src = np.stack([rng.uniform(0, 360, M), rng.uniform(-60, 80, M), rng.uniform(0.8, 2.0, M)], axis=1)
    return ir, src


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("record_m", [False, True])
def test_reader_matches_an_independent_netcdf_implementation(tmp_path, version, record_m):
    rng = np.random.default_rng(100 + version + 10 * record_m)
    ir, src = make_set(rng)
    M, R, N = ir.shape
    delay = rng.uniform(0, 30, (M, R))
    path = str(tmp_path / "set.sofa")
    write_sofa(path, ir, delay, 48000.0, src, version=version, record_m=record_m)
    ref = netcdf_file(path, "r", mmap=False)  # scipy's reader: the independent statement of the container
    s = bbx.SOFA(path)
    assert (s.num_measurements, s.num_receivers, s.num_emitters, s.ir_length) == (M, R, 1, N)
    assert s.get_samplerate() == float(ref.variables["Data.SamplingRate"][0]) == 48000.0
    for m in range(M):
        for r in range(R):
            want = np.asarray(ref.variables["Data.IR"][m, r], dtype=np.float64).astype(np.float32)  # double -> float: one rounding
            assert np.array_equal(s.get_ir(m, r), want)
            assert s.get_delay(m, r) == float(ref.variables["Data.Delay"][m, r])
        p, sph = s.get_position(bbx.SOFA.SOURCE, m)
        assert sph and np.array_equal(p, np.asarray(ref.variables["SourcePosition"][m], dtype=np.float64))
    p, sph = s.get_position(bbx.SOFA.LISTENER, M - 1)  # [I][C]: the one row serves every measurement
    assert not sph and np.array_equal(p, [0.0, 0.0, 0.0])
    p, sph = s.get_position(bbx.SOFA.RECEIVER, 1)
    assert not sph and np.allclose(p, [0.0, -0.09, 0.0])
    assert s.get_attribute("SOFAConventions") == "SimpleFreeFieldHRIR"
    assert s.get_attribute("DataType") == "FIR"
    # truncation / zero padding of get_ir
    assert np.array_equal(s.get_ir(2, 1, n=8), s.get_ir(2, 1)[:8])
    padded = s.get_ir(2, 1, n=N + 5)
    assert np.array_equal(padded[:N], s.get_ir(2, 1)) and not padded[N:].any()
    s.close()
    ref.close()


def test_fire_sets_float_data_and_shared_delay(tmp_path):
    rng = np.random.default_rng(5)
    ir, src = make_set(rng, M=3, R=2, N=16, E=4)
    delay = rng.uniform(0, 4, (1, 2, 4))
    path = str(tmp_path / "fire.sofa")
    write_sofa(path, ir.astype(np.float32), delay, 44100.0, src, fire=True, ir_dtype="f")
    s = bbx.SOFA(path)
    assert (s.num_measurements, s.num_receivers, s.num_emitters, s.ir_length) == (3, 2, 4, 16)
    for m in range(3):
        for r in range(2):
            for e in range(4):
                assert np.array_equal(s.get_ir(m, r, e), ir[m, r, e].astype(np.float32))
                assert s.get_delay(m, r, e) == delay[0, r, e]
    assert s.get_samplerate(2) == 44100.0
    s.close()


def test_open_from_memory_equals_open_from_file(tmp_path):
    rng = np.random.default_rng(6)
    ir, src = make_set(rng, M=4)
    path = str(tmp_path / "m.sofa")
    write_sofa(path, ir, np.zeros((1, 2)), 48000.0, src, extra_vars=False)
    a, b = bbx.SOFA(path), bbx.SOFA(data=open(path, "rb").read())
    for m in range(4):
        assert np.array_equal(a.get_ir(m, 1), b.get_ir(m, 1))
        assert a.get_delay(m, 0) == b.get_delay(m, 0) == 0.0
    with pytest.raises(bbx.BbxError, match="no such variable"):
        a.get_position(bbx.SOFA.LISTENER, 0)


@pytest.mark.parametrize("src_type", ["spherical", "cartesian"])
def test_nearest_measurement_equals_brute_force(tmp_path, src_type):
    rng = np.random.default_rng(8)
    M = 200
    ir, src = make_set(rng, M=M, N=4)
    if src_type == "cartesian":
        src = rng.uniform(-2, 2, (M, 3))
    path = str(tmp_path / "n.sofa")
    write_sofa(path, ir, np.zeros((1, 2)), 48000.0, src, src_type=src_type, extra_vars=False)
    s = bbx.SOFA(path)

    def cart(p, sph):
        p = np.asarray(p, dtype=np.float64)
        if not sph:
            return p
        az, el, r = np.deg2rad(p[..., 0]), np.deg2rad(p[..., 1]), p[..., 2]
        return np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], axis=-1)

    cs = cart(src, src_type == "spherical")
    for _ in range(50):
        q = np.array([rng.uniform(0, 360), rng.uniform(-90, 90), rng.uniform(0.5, 2.5)])
        want = int(np.argmin(((cs - cart(q, True)) ** 2).sum(axis=1)))
        assert s.nearest_measurement(q, spherical=True) == want
        # direction-only query (radius 0): both sides normalised
        qd = np.array([q[0], q[1], 0.0])
        unit = cs / np.linalg.norm(cs, axis=1, keepdims=True)
        want = int(np.argmin(((unit - cart([q[0], q[1], 1.0], True)) ** 2).sum(axis=1)))
        assert s.nearest_measurement(qd, spherical=True) == want
    # an exact hit returns that measurement
    for m in (0, 17, M - 1):
        assert s.nearest_measurement(src[m], spherical=(src_type == "spherical")) == m


def test_rejects_what_it_cannot_read(tmp_path):
    rng = np.random.default_rng(9)
    ir, src = make_set(rng, M=3)
    path = str(tmp_path / "ok.sofa")
    write_sofa(path, ir, np.zeros((1, 2)), 48000.0, src)
    blob = open(path, "rb").read()
    with pytest.raises(bbx.BbxError, match="HDF5"):
        bbx.SOFA(data=b"\x89HDF\r\n\x1a\n" + bytes(64))
    with pytest.raises(bbx.BbxError, match="not a netCDF"):
        bbx.SOFA(data=b"RIFF" + bytes(64))
    with pytest.raises(bbx.BbxError, match="version 5"):
        bbx.SOFA(data=b"CDF\x05" + blob[4:])
    with pytest.raises(bbx.BbxError):
        bbx.SOFA(data=blob[: len(blob) // 2])  # data runs past the end
    with pytest.raises(bbx.BbxError):
        bbx.SOFA(data=blob[:40])  # truncated header
    with pytest.raises(bbx.BbxError, match="cannot open"):
        bbx.SOFA(str(tmp_path / "absent.sofa"))
    # a netCDF file that is not a SOFA impulse-response set
    f = netcdf_file(str(tmp_path / "plain.nc"), "w")
    f.createDimension("x", 3)
    v = f.createVariable("v", "d", ("x",))
    v[:] = [1, 2, 3]
    f.close()
    with pytest.raises(bbx.BbxError, match="Conventions"):
        bbx.SOFA(str(tmp_path / "plain.nc"))
    s = bbx.SOFA(path)
    with pytest.raises(bbx.BbxError, match="outside"):
        s.get_ir(3, 0)
    with pytest.raises(bbx.BbxError, match="no global attribute"):
        s.get_attribute("Nope")
    # every header byte flipped in turn: an error or a handle, never a crash
    for i in range(4, min(len(blob), 700), 3):
        bad = bytearray(blob)
        bad[i] ^= 0xFF
        try:
            bbx.SOFA(data=bytes(bad)).close()
        except bbx.BbxError:
            pass


@pytest.mark.gpu
def test_filter_bank_from_a_sofa_set_renders_like_direct_convolution(tmp_path):
    """A binaural renderer's use of the set: one bank per ear, nearest-measurement selection with the set's delays."""
    rng = np.random.default_rng(11)
    M, R, N, B = 12, 2, 300, 128
    ir, src = make_set(rng, M=M, R=R, N=N)
    ir /= np.sqrt((ir ** 2).sum(axis=-1, keepdims=True))
    delay = np.round(rng.uniform(0, 20, (M, R)))
    path = str(tmp_path / "hrir.sofa")
    write_sofa(path, ir, delay, 48000.0, src)
    s = bbx.SOFA(path)
    eng = bbx.Convolver(B, 3, 1, n_outputs=R, n_paths=R, mode=bbx.MODE_ROUTED, max_blocks=4, max_delay=32)
    banks = [s.create_filters(eng, r) for r in range(R)]
    assert all(len(b) == M and b[0].partitions == 3 for b in banks)
    q = src[5] + np.array([0.2, -0.1, 0.0])
    m = s.nearest_measurement(q)
    assert m == 5
    for r in range(R):
        eng.SetRoute(r, 0, r, 1.0)
        eng.SelectFilter(r, banks[r][m], delay=s.get_delay(m, r))
    T = 4
    x = rng.uniform(-1, 1, T * B).astype(np.float32)
    y = eng.Convolve(x, bbx.FMT_FLOAT, 1, bbx.FMT_FLOAT, R, T * B).view(np.float32).reshape(T * B, R)
    for r in range(R):
        h = ir[m, r].astype(np.float32).astype(np.float64)
        want = np.convolve(x.astype(np.float64), h)[: T * B]
        d = int(delay[m, r])
        want = np.concatenate([np.zeros(d), want])[: T * B]
        err = y[:, r] - want
        snr = 10 * np.log10((want ** 2).sum() / max((err ** 2).sum(), 1e-300))
        assert snr >= 110.0 and np.abs(err).max() <= 1e-5 * np.abs(want).max(), (r, snr)
    with pytest.raises(bbx.BbxError, match="receiver"):
        s.create_filters(eng, R)
    eng.close()
