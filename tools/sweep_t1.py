"""T = 1 latency of the C3 engine against the L2-residency fraction of the streaming MAC (mac_l2_keep_16ths).
    python tools/sweep_t1.py [calls]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bbcat_dsp_b200 as bbx  # noqa: E402

ncalls = int(sys.argv[1]) if len(sys.argv) > 1 else 400
nch, L, B = 128, 144000, 512
P = (L + B - 1) // B
rng = np.random.default_rng(1)
irs = [(rng.standard_normal(L) * np.exp(-6.9 * np.arange(L) / L)).astype(np.float32) for _ in range(nch)]
hin, hout = bbx.PinnedBuffer(B * nch * 4), bbx.PinnedBuffer(B * nch * 4)
hin.array[:] = rng.uniform(-1, 1, B * nch).astype(np.float32).view(np.uint8)
for keep in (3, 0, 17, 2, 4, 5, 6, 8):   # 0 = default (3), 17 = hints off
    eng = bbx.Convolver(B, P, nch, max_blocks=1, mac_l2_keep_16ths=keep)
    for c in range(nch):
        eng.SelectFilter(c, eng.CreateFilter(irs[c]))
    lat = []
    eng.profile_mac(True)
    for i in range(ncalls):
        t0 = time.perf_counter()
        eng.ConvolveHostPtr(hin.ptr, bbx.FMT_FLOAT, nch, hout.ptr, bbx.FMT_FLOAT, nch, B)
        lat.append(1e6 * (time.perf_counter() - t0))
    mac = eng.mac_time()
    lat = np.array(lat[50:])
    print("l2_keep_16ths %2d: p50 %.1f us  p99 %.1f us  MAC kernel %.1f us per call" % (
        keep, np.percentile(lat, 50), np.percentile(lat, 99), 1e3 * mac["ms"] / mac["launches"]), flush=True)
    eng.close()
