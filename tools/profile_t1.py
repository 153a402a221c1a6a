"""T = 1 streaming calls of the C3 engine (pinned buffers) for an ncu launch list / timeline of the latency path.
    python tools/profile_t1.py [calls] [channels] [taps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bbcat_dsp_b200 as bbx  # noqa: E402

ncalls = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
L = int(sys.argv[3]) if len(sys.argv) > 3 else 144000
B = 512
P = (L + B - 1) // B
eng = bbx.Convolver(B, P, nch, max_blocks=1)
rng = np.random.default_rng(1)
for c in range(nch):
    h = (rng.standard_normal(L) * np.exp(-6.9 * np.arange(L) / L)).astype(np.float32)
    eng.SelectFilter(c, eng.CreateFilter(h))
hin, hout = bbx.PinnedBuffer(B * nch * 4), bbx.PinnedBuffer(B * nch * 4)
hin.array[:] = rng.uniform(-1, 1, B * nch).astype(np.float32).view(np.uint8)
lat = []
for i in range(ncalls):
    t0 = time.perf_counter()
    eng.ConvolveHostPtr(hin.ptr, bbx.FMT_FLOAT, nch, hout.ptr, bbx.FMT_FLOAT, nch, B)
    lat.append(1e6 * (time.perf_counter() - t0))
print("direct calls", eng.direct_calls(), "median us", float(np.median(lat[len(lat) // 2:])))
eng.close()
