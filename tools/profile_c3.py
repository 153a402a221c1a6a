"""C3 workload (128 ch x 144000 taps, B = 512, 64-block calls) for ncu captures: argv[1] = calls (default 3)."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbcat_dsp_b200 as bbx
B, nch, T, L = 512, int(os.environ.get("C3_CH", "128")), 64, 144000
P = (L + B - 1) // B
eng = bbx.Convolver(B, P, nch, max_blocks=T)
rng = np.random.default_rng(1)
h = (rng.standard_normal(L) * np.exp(-6.9 * np.arange(L) / L)).astype(np.float32)
for c in range(nch):
    eng.SelectFilter(c, eng.CreateFilter(np.roll(h, c)))
x = rng.uniform(-1, 1, (T * B, nch)).astype(np.float32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    y = eng.Convolve(x, bbx.FMT_FLOAT, nch, bbx.FMT_FLOAT, nch, T * B)
eng.close()
