#!/usr/bin/env python
"""Host <-> device copies against the C3 step: who slows whom?

Runs, on one GPU, (a) the device-resident C3 step alone, (b) plain pinned copies of one step's input and output alone
(two streams, both directions at once), (c) both at the same time, and prints the per-step compute time and the per-copy
time of each case.  The MAC kernel is chosen by the engine's usual knobs (BBX_TBW, BBX_TBS_*, --tile).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    import torch
    import bbcat_dsp_b200 as bbx
    import bench

    B, P, NCH, L, T = bench.B, bench.P, bench.NCH, bench.L, bench.T
    eng = bbx.Convolver(B, P, NCH, max_blocks=T, device=0, mac_time_tile=args.tile)
    for c in range(NCH):
        eng.SelectFilter(c, eng.CreateFilter(bench.make_ir(2000 + c, L)))
    frames = T * B
    x = (torch.rand((frames, NCH), device="cuda") * 2 - 1).contiguous()
    y = torch.empty_like(x)
    nbytes = frames * NCH * 4
    hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def compute(n):
        for _ in range(n):
            eng.ConvolveDev(x.data_ptr(), bbx.FMT_FLOAT, NCH, y.data_ptr(), bbx.FMT_FLOAT, NCH, frames)

    def copies(n):
        evs = []
        for s, fn in ((s_in, lambda: din.copy_(hin, non_blocking=True)), (s_out, lambda: hout.copy_(dout, non_blocking=True))):
            with torch.cuda.stream(s):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    fn()
                e1.record()
                evs.append((e0, e1))
        return evs

    n = args.steps
    compute(10)
    eng.Sync()
    torch.cuda.synchronize()
    res = {}
    # (a) compute alone
    eng.profile_mac(True)
    eng.timer_start()
    compute(n)
    ms = eng.timer_stop()
    mac = eng.mac_time()
    eng.profile_mac(False)
    res["compute_alone_ms"] = ms / n
    res["mac_alone_ms"] = mac["ms"] / max(1, mac["launches"])
    res["kernel"] = eng.mac_kernel_name()
    # (b) copies alone
    torch.cuda.synchronize()
    evs = copies(n)
    torch.cuda.synchronize()
    res["h2d_alone_ms"], res["d2h_alone_ms"] = [e0.elapsed_time(e1) / n for e0, e1 in evs]
    # (c) both
    torch.cuda.synchronize()
    eng.profile_mac(True)
    evs = copies(n)
    eng.timer_start()
    compute(n)
    ms = eng.timer_stop()
    torch.cuda.synchronize()
    mac = eng.mac_time()
    eng.profile_mac(False)
    res["compute_with_copies_ms"] = ms / n
    res["mac_with_copies_ms"] = mac["ms"] / max(1, mac["launches"])
    res["h2d_with_compute_ms"], res["d2h_with_compute_ms"] = [e0.elapsed_time(e1) / n for e0, e1 in evs]
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))


if __name__ == "__main__":
    main()
