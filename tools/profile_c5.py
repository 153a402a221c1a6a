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
if len(sys.argv) > 2 and sys.argv[2] == "trace":
    eng.tensor_trace(128, enable=True)
    y = eng.Convolve(x, bbx.FMT_FLOAT, nin, bbx.FMT_FLOAT, nout, T * B)
    tr = eng.tensor_trace(128).astype(np.float64)
    names = ["loader total", "loader wait raw slot", "mma total", "mma wait operands", "mma wait read-out", "epi total",
             "epi wait acc"] + ["prod%d %s" % (g, w) for g in range(3) for w in ("total", "wait raw", "wait stage")]
    if len(sys.argv) > 3 and sys.argv[3] == "fine":   # libbbx built with -DBBX_TC_FINE_TRACE: group 0 only, split by phase
        names = names[:10] + ["prod0 A part", "prod0 B part", "prod0 wait::st", "prod0 fences+arrive", "loader prefetch issue", "loader copy issue"]
    for k, nme in enumerate(names):
        print("%-22s mean %9.0f  min %9.0f  max %9.0f cycles" % (nme, tr[:, k].mean(), tr[:, k].min(), tr[:, k].max()))
eng.close()
