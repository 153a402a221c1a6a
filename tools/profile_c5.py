import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import bbcat_dsp_b200 as bbx
B, nin, nout, T = 512, 64, 64, 64
eng = bbx.Convolver(B, 8, nin, n_outputs=nout, mode=bbx.MODE_MIMO, max_blocks=T)
rng = np.random.default_rng(1)
for o in range(nout):
    for i in range(nin):
        h = (rng.standard_normal(4096) * np.exp(-6.9 * np.arange(4096) / 4096)).astype(np.float32)
        eng.SelectFilter(o * nin + i, eng.CreateFilter(h))
x = rng.uniform(-1, 1, (T * B, nin)).astype(np.float32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    y = eng.Convolve(x, bbx.FMT_FLOAT, nin, bbx.FMT_FLOAT, nout, T * B)
print(eng.tensor_status())
eng.close()
