#!/usr/bin/env python
"""The standalone entry points of the path (TransferSamples, MixSamples, FractionalSample) on device-resident buffers,
next to the single-threaded CPU rates bench.py reports in cpu_baseline.reference_functions.  CUDA events on
the null stream the *_dev entry points are given; prints one JSON object.

    python tools/bench_functions.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbcat_dsp_b200 as bbx  # noqa: E402


def main():
    import torch
    lib, vp = bbx.lib(), C.c_void_p
    nch, nfr = 32, 1 << 20   # 32 channels x 1 Mi frames (21.8 s at 48 kHz)
    x = (torch.rand(nch * nfr, device="cuda") * 2 - 1).contiguous()
    s24 = torch.zeros(nch * nfr * 3, dtype=torch.uint8, device="cuda")
    back = torch.zeros(nch * nfr, device="cuda")
    bus = torch.zeros(2 * nfr, device="cuda")
    ring = x[:4096].clone()
    npos = 1 << 22
    pos = (torch.rand(npos, device="cuda", dtype=torch.float64) * 3900 + 50).contiguous()
    outp = torch.zeros(npos, device="cuda", dtype=torch.float64)

    def timed(fn, samples, reps=20):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return samples * reps / (e0.elapsed_time(e1) * 1e-3) / 1e6

    def chk(rc):
        if rc:
            raise RuntimeError(lib.bbx_last_error().decode())

    st = vp(0)
    res = {"unit": "Msample/s", "geometry": "%d channels x %d frames, device-resident" % (nch, nfr)}
    res["transfer_f32_to_s24_32ch"] = timed(lambda: chk(lib.bbx_transfer_samples_dev(vp(x.data_ptr()), 4, 0, 0, nch, vp(s24.data_ptr()), 2, 0, 0,
                                                                                      nch, nch, nfr, st)), nch * nfr)
    res["transfer_s24_to_f32_32ch"] = timed(lambda: chk(lib.bbx_transfer_samples_dev(vp(s24.data_ptr()), 2, 0, 0, nch, vp(back.data_ptr()), 4, 0,
                                                                                      0, nch, nch, nfr, st)), nch * nfr)

    def mix32():
        for p in range(32):
            chk(lib.bbx_mix_samples_f32_dev(vp(x.data_ptr() + 4 * p * nfr), 0, 1, vp(bus.data_ptr()), p & 1, 2, 1, nfr, C.c_float(0.5), st))
    res["mix_32_mono_paths_to_stereo"] = timed(mix32, 32 * nfr, reps=5)
    res["fractional_sample_f32"] = timed(lambda: chk(lib.bbx_fractional_samples_f32_dev(vp(ring.data_ptr()), 0, 1, 4096, vp(pos.data_ptr()), npos,
                                                                                         vp(outp.data_ptr()), st)), npos)
    # BiQuadFilterBank: 8 filters on 1024 channels x 16 Ki frames -- the fused pass against the same filters run one bank
    # after the other (the reference's loop, one kernel per filter)
    fch, ffr, nf = 1024, 1 << 14, 8
    xs = (torch.rand(fch * ffr, device="cuda") * 2 - 1).contiguous()
    ys = torch.zeros(fch * ffr, device="cuda")
    fb = bbx.BiQuadFilterBank(fch, nf)
    singles = [bbx.BiQuadBank(fch) for _ in range(nf)]
    for f in range(nf):
        c = bbx.BiQuadCalcCoeffs(7, 300.0 * (f + 1), 48000.0, 3.0, 1.0)
        fb.SetCoeffs(f, c)
        singles[f].SetCoeffs(c)
    res["fbank_8_filters_fused_1024ch"] = timed(lambda: fb.ProcessDev(vp(xs.data_ptr()), vp(ys.data_ptr()), fch, fch, fch, ffr, st), fch * ffr, reps=5)

    def chain():
        chk(lib.bbx_biquad_process_dev(singles[0].h, vp(xs.data_ptr()), vp(ys.data_ptr()), fch, fch, fch, ffr, st))
        for q in singles[1:]:
            chk(lib.bbx_biquad_process_dev(q.h, vp(ys.data_ptr()), vp(ys.data_ptr()), fch, fch, fch, ffr, st))
    res["fbank_8_filters_one_pass_each_1024ch"] = timed(chain, fch * ffr, reps=5)
    print(json.dumps({"gpu_functions": res}))


if __name__ == "__main__":
    main()
