"""T = 1 streaming calls of the short-filter configs (C1 / C2 / C4) for an ncu launch list of the latency path, plus the host
round trip of an empty-ish call for reference.
    python tools/profile_lat.py C1|C2|C4 [calls] [fused 0/1]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bbcat_dsp_b200 as bbx  # noqa: E402


def make_ir(seed, n):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal(n) * np.exp(-6.9 * np.arange(n) / n)
    return (h / np.sqrt((h ** 2).sum())).astype(np.float32)


cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
ncalls = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fused = int(sys.argv[3]) if len(sys.argv) > 3 else 1
if cfg == "C1":
    B, nin, nout, fi, fo = 1024, 2, 2, 4, 4
    eng = bbx.Convolver(B, 8, 2, max_blocks=1)
    for c in range(2):
        eng.SelectFilter(c, eng.CreateFilter(make_ir(2000 + c, 8192)))
elif cfg == "C2":
    B, nin, nout, fi, fo = 256, 64, 2, 4, 4
    eng = bbx.Convolver(B, 2, 64, n_outputs=2, n_paths=128, mode=bbx.MODE_ROUTED, max_blocks=1, max_delay=48)
    for s in range(64):
        for ear in range(2):
            p = 2 * s + ear
            eng.SetRoute(p, s, ear, 1.0 / 8)
            eng.SelectFilter(p, eng.CreateFilter(make_ir(2000 + p, 512)), delay=float((s * (1 + ear)) % 40))
else:
    B, nin, nout, fi, fo = 512, 32, 32, 2, 2
    eng = bbx.Convolver(B, 8, 32, max_blocks=1, max_delay=64, fractional_delay=True)
    for c in range(32):
        eng.SelectFilter(c, eng.CreateFilter(make_ir(2000 + c, 4096)), delay=20.5)
eng.set_fused(bool(fused))
hin = bbx.PinnedBuffer(B * nin * bbx.FMT_BYTES[fi])
hout = bbx.PinnedBuffer(B * nout * bbx.FMT_BYTES[fo])
hin.array[:] = np.random.default_rng(5).integers(0, 255, hin.nbytes, dtype=np.uint8) if fi < 4 else \
    np.random.default_rng(5).uniform(-1, 1, B * nin).astype(np.float32).view(np.uint8)
lat = []
for i in range(ncalls):
    t0 = time.perf_counter()
    eng.ConvolveHostPtr(hin.ptr, fi, nin, hout.ptr, fo, nout, B)
    lat.append(1e6 * (time.perf_counter() - t0))
lat = np.array(lat[len(lat) // 4:])
print(cfg, "fused", fused, "fused calls", eng.fused_calls(), "direct calls", eng.direct_calls(), "p50 us %.1f p99 %.1f min %.1f" % (
    np.percentile(lat, 50), np.percentile(lat, 99), lat.min()))
eng.close()
