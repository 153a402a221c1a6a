#!/usr/bin/env python
"""Timeline of the host-buffer pipeline of the C3 step, rebuilt in Python around bbx_process_dev.

The same three streams and events as bbx_process_async (H2D copy stream, engine stream, D2H copy stream; --slots staging
buffers), with a CUDA event at the start and end of every copy and every step, so that the steady state can be read as a
timeline: which edge of the dependency graph a step waits on.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--slots", type=int, default=2)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--show", type=int, default=6)
    ap.add_argument("--engine", action="store_true", help="trace bbx_process_async itself (bbx_engine_io_trace) instead of the Python rebuild")
    args = ap.parse_args()
    import torch
    import bbcat_dsp_b200 as bbx
    import bench

    B, NCH, L, T = bench.B, bench.NCH, bench.L, bench.T
    eng = bbx.Convolver(B, bench.P, NCH, max_blocks=T, device=0, mac_time_tile=args.tile)
    for c in range(NCH):
        eng.SelectFilter(c, eng.CreateFilter(bench.make_ir(2000 + c, L)))
    frames = T * B
    nbytes = frames * NCH * 4
    if args.engine:
        hb = [bbx.PinnedBuffer(nbytes) for _ in range(4)]
        for h in hb[:2]:
            h.array.view("float32")[:] = 0.25
        n = args.steps
        for i in range(8):
            eng.ConvolveHostPtrAsync(hb[i & 1].ptr, bbx.FMT_FLOAT, NCH, hb[2 + (i & 1)].ptr, bbx.FMT_FLOAT, NCH, frames)
        eng.Sync()
        eng.io_trace(n)
        for i in range(n):
            eng.ConvolveHostPtrAsync(hb[i & 1].ptr, bbx.FMT_FLOAT, NCH, hb[2 + (i & 1)].ptr, bbx.FMT_FLOAT, NCH, frames)
        tr = eng.io_trace_read()
        period = (tr[n - 1][5] - tr[n // 2][5]) / (n - 1 - n // 2)
        dur = lambda a, b: sum(r[b] - r[a] for r in tr[n // 2:]) / (n - n // 2)
        print(json.dumps({"kernel": eng.mac_kernel_name(), "pipeline": "bbx_process_async", "period_ms": round(period, 4),
                          "h2d_ms": round(dur(0, 1), 4), "compute_ms": round(dur(2, 3), 4), "d2h_ms": round(dur(4, 5), 4)}))
        base = tr[n - args.show][0]
        for i in range(n - args.show, n):
            print("step %3d  h2d %.3f-%.3f  comp %.3f-%.3f  d2h %.3f-%.3f" % ((i,) + tuple(v - base for v in tr[i])))
        return
    S = args.slots
    hin = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    hout = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for h in hin:
        h.view(torch.float32).uniform_(-1, 1)
    sin = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(S)]
    sout = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(S)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    s_eng = torch.cuda.ExternalStream(bbx._lib.bbx_engine_get_stream(eng.h))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    n = args.steps
    E = {k: [ev() for _ in range(n)] for k in ("h0", "h1", "c0", "c1", "d0", "d1")}
    torch.cuda.synchronize()
    t0 = ev()
    t0.record(s_eng)
    for i in range(n):
        k = i % S
        with torch.cuda.stream(s_in):
            if i >= S:
                s_in.wait_event(E["c1"][i - S])
            E["h0"][i].record()
            sin[k].copy_(hin[i & 1], non_blocking=True)
            E["h1"][i].record()
        with torch.cuda.stream(s_eng):
            s_eng.wait_event(E["h1"][i])
            if i >= S:
                s_eng.wait_event(E["d1"][i - S])
            E["c0"][i].record()
            eng.ConvolveDev(sin[k].data_ptr(), bbx.FMT_FLOAT, NCH, sout[k].data_ptr(), bbx.FMT_FLOAT, NCH, frames)
            E["c1"][i].record()
        with torch.cuda.stream(s_out):
            s_out.wait_event(E["c1"][i])
            E["d0"][i].record()
            hout[i & 1].copy_(sout[k], non_blocking=True)
            E["d1"][i].record()
    torch.cuda.synchronize()
    tl = {k: [t0.elapsed_time(e) for e in v] for k, v in E.items()}
    period = (tl["d1"][n - 1] - tl["d1"][n // 2]) / (n - 1 - n // 2)
    dur = lambda a, b: sum(tl[b][i] - tl[a][i] for i in range(n // 2, n)) / (n - n // 2)
    print(json.dumps({"kernel": eng.mac_kernel_name(), "slots": S, "period_ms": round(period, 4), "h2d_ms": round(dur("h0", "h1"), 4),
                      "compute_ms": round(dur("c0", "c1"), 4), "d2h_ms": round(dur("d0", "d1"), 4)}))
    base = tl["h0"][n - args.show]
    for i in range(n - args.show, n):
        print("step %3d  h2d %.3f-%.3f  comp %.3f-%.3f  d2h %.3f-%.3f" % ((i,) + tuple(tl[k][i] - base for k in ("h0", "h1", "c0", "c1", "d0", "d1"))))


if __name__ == "__main__":
    main()
