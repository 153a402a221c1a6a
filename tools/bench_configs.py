#!/usr/bin/env python
"""Secondary measurements: every BASELINE.json config on one B200 (bench.py covers C3, the headline).

For each config: throughput in channel-seconds of audio per second with inputs resident in HBM (CUDA events on
the engine streams) and the per-block latency of the streaming call (T = 1, bbx_process with pinned host buffers,
host clock), 1000 blocks after 100 warm-up.  Prints one JSON object per config; results are copied into DESIGN.md.

    python tools/bench_configs.py [--configs C1,C2,C4,C5] [--steps 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bbcat_dsp_b200 as bbx  # noqa: E402

FS = 48000


def make_ir(seed, n):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal(n) * np.exp(-6.9 * np.arange(n) / n)
    h /= np.sqrt((h ** 2).sum())
    return h.astype(np.float32)


def latency(eng, fmt_in, in_ch, fmt_out, out_ch, B, n=1000, warm=100):
    hin = bbx.PinnedBuffer(B * in_ch * bbx.FMT_BYTES[fmt_in])
    hout = bbx.PinnedBuffer(B * out_ch * bbx.FMT_BYTES[fmt_out])
    hin.array[:] = np.random.default_rng(5).integers(0, 255, hin.nbytes, dtype=np.uint8) if fmt_in < 4 else \
        np.random.default_rng(5).uniform(-1, 1, B * in_ch).astype(np.float32).view(np.uint8)
    lat = []
    for i in range(n + warm):
        t0 = time.perf_counter()
        eng.ConvolveHostPtr(hin.ptr, fmt_in, in_ch, hout.ptr, fmt_out, out_ch, B)
        if i >= warm:
            lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e6
    return {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "blocks": n,
            "block_period_us": 1e6 * B / FS}


def throughput(eng, fmt_in, in_ch, fmt_out, out_ch, B, T, steps, channels, before_step=None):
    import torch
    frames = T * B
    x = torch.randint(0, 255, (frames * in_ch * bbx.FMT_BYTES[fmt_in],), dtype=torch.uint8, device="cuda") if fmt_in < 4 \
        else (torch.rand(frames * in_ch, device="cuda") * 2 - 1).view(torch.uint8)
    y = torch.empty(frames * out_ch * bbx.FMT_BYTES[fmt_out], dtype=torch.uint8, device="cuda")

    def step(i):
        n = frames
        if before_step:
            n = before_step(i) * B
        eng.ConvolveDev(x.data_ptr(), fmt_in, in_ch, y.data_ptr(), fmt_out, out_ch, n)
        return n

    for i in range(5):
        step(i)
    eng.Sync()
    l0 = eng.launch_count()
    eng.timer_start()
    total = 0
    for i in range(steps):
        total += step(i + 5)
    ms = eng.timer_stop()
    return {"channel_s_per_s": channels * total / FS / (ms * 1e-3), "ms_per_step": ms / steps, "blocks_per_step": T,
            "launches_per_step": (eng.launch_count() - l0) / steps, "x_realtime": total / FS / (ms * 1e-3)}


def c1(steps):
    B, L, nch, T = 1024, 8192, 2, 64
    eng = bbx.Convolver(B, 8, nch, max_blocks=T)
    for c in range(nch):
        eng.SelectFilter(c, eng.CreateFilter(make_ir(2000 + c, L)))
    r = {"config": "C1 stereo 2ch, 8192 taps, B=1024, f32", "channels": nch}
    r.update(throughput(eng, 4, nch, 4, nch, B, T, steps, nch))
    r["latency"] = latency(eng, 4, nch, 4, nch, B)
    eng.close()
    return r


def c2(steps):
    B, L, nsrc, T = 256, 512, 64, 64
    eng = bbx.Convolver(B, 2, nsrc, n_outputs=2, n_paths=2 * nsrc, mode=bbx.MODE_ROUTED, max_blocks=T, max_delay=48)
    for s in range(nsrc):
        for ear in range(2):
            p = 2 * s + ear
            eng.SetRoute(p, s, ear, 1.0 / 8)
            eng.SelectFilter(p, eng.CreateFilter(make_ir(2000 + p, L)), delay=float((s * (1 + ear)) % 40))
    r = {"config": "C2 binaural 64 sources x 2 ears (128 paths), 512-tap HRIRs, B=256, ITD delays, f32", "channels": nsrc}
    r.update(throughput(eng, 4, nsrc, 4, 2, B, T, steps, nsrc))
    r["latency"] = latency(eng, 4, nsrc, 4, 2, B)
    eng.close()
    return r


def c4(steps):
    B, L, nch, nbank = 512, 4096, 32, 16
    eng = bbx.Convolver(B, 8, nch, max_blocks=10, max_delay=64, fractional_delay=True)
    bank = [[eng.CreateFilter(make_ir(2000 + 16 * c + k, L)) for k in range(nbank)] for c in range(nch)]
    state = {"m": 0}

    def before(i):
        m = state["m"]
        state["m"] += 1
        eng.SelectFilters(range(nch), [bank[c][(m + c) % nbank] for c in range(nch)],
                          delays=[16 + 37.3 * ((m * 7 + c) % 11) / 11 for c in range(nch)], crossfade=[m > 0] * nch)
        # switch every 100 ms: block index ceil(m * 4800 / 512)
        cd = lambda a: -(-a // 512)
        return cd((m + 1) * 4800) - cd(m * 4800)

    r = {"config": "C4 32ch dynamic IR: bank of 16 IRs/ch (4096 taps), select every 100 ms with crossfade + fractional "
                   "delay, B=512, s24 in/out", "channels": nch}
    r.update(throughput(eng, 2, nch, 2, nch, B, 10, steps, nch, before_step=before))
    for c in range(nch):
        eng.SelectFilter(c, bank[c][0], delay=20.5)
    r["latency"] = latency(eng, 2, nch, 2, nch, B)
    eng.close()
    return r


def c5(steps):
    B, L, nin, nout, T = 512, 4096, 64, 64, 64
    eng = bbx.Convolver(B, 8, nin, n_outputs=nout, mode=bbx.MODE_MIMO, max_blocks=T)
    for o in range(nout):
        for i in range(nin):
            eng.SelectFilter(o * nin + i, eng.CreateFilter(make_ir(2000 + 64 * o + i, L)))
    r = {"config": "C5 MIMO 64 in x 64 out, 4096-tap matrix (4096 paths), B=512, f32, single GPU", "channels": nout}
    r.update(throughput(eng, 4, nin, 4, nout, B, T, steps, nout))
    # HBM roofline of the streaming MAC: 16 * P * K bytes of spectra per path-block, FDL shared
    bytes_per_block = 16 * 8 * 513 * nin * nout / 2 + 16 * 513 * (nin + nout)
    r["hbm_algorithmic_GBps"] = bytes_per_block * T / (r["ms_per_step"] * 1e-3) / 1e9
    r["latency"] = latency(eng, 4, nin, 4, nout, B)
    eng.close()
    return r


def c5_sharded(steps, peer=False):
    """C5 input-sharded over the ranks of a torchrun launch (one process per GPU): rank g holds 64 / world inputs and
    all 64 outputs; per call one ncclReduceScatter of the partial output spectra (16.8 MB at T = 64); rank g converts
    64 / world outputs.  Device time = max over ranks between barriers; prints on rank 0.

        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
            tools/bench_configs.py --configs C5S
    """
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    B, L, nin, nout, T = 512, 4096, 64, 64, 64
    uid = [bbx.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    comm = bbx.Comm(world, rank, uid[0], device=local)
    i0, ni = bbx.shard_range(nin, rank, world)
    eng = bbx.Convolver(B, 8, ni, n_outputs=nout, mode=bbx.MODE_MIMO, max_blocks=T, mimo_shard_world=world,
                        mimo_shard_rank=rank, device=local)
    if peer and world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, eng.PeerExport())
        eng.PeerAttach(handles)
    else:
        eng.SetComm(comm)
    for o in range(nout):
        for i in range(ni):
            eng.SelectFilter(o * ni + i, eng.CreateFilter(make_ir(2000 + 64 * o + i0 + i, L)))
    nloc = nout // world
    frames = T * B
    x = (torch.rand(frames * ni, device="cuda") * 2 - 1).view(torch.uint8)
    y = torch.empty(frames * nloc * 4, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        eng.ConvolveDev(x.data_ptr(), 4, ni, y.data_ptr(), 4, nloc, frames)
    eng.Sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.timer_start()
    for _ in range(steps):
        eng.ConvolveDev(x.data_ptr(), 4, ni, y.data_ptr(), 4, nloc, frames)
    ms = eng.timer_stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    r = {"config": "C5 MIMO 64 x 64, 4096 taps, B=512, f32, input-sharded over %d GPU(s), %s of %d MB per "
                   "step" % (world, "peer-memory mixdown (NVLink stores + epoch flags)" if peer and world > 1 else "ncclReduceScatter",
                             nout * T * B * 8 >> 20), "channels": nout, "n_gpus": world,
         "channel_s_per_s": nout * steps * frames / FS / (ms * 1e-3), "ms_per_step": ms / steps, "blocks_per_step": T}
    if world > 1:
        dist.barrier()  # peer mode: nobody frees a receive buffer the others still have mapped
    eng.close()
    comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return r if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C4,C5")
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    fns = {"C1": c1, "C2": c2, "C4": c4, "C5": c5, "C5S": c5_sharded, "C5P": lambda steps: c5_sharded(steps, peer=True)}
    for name in args.configs.split(","):
        r = fns[name](args.steps)
        if r is not None:
            print(json.dumps({name: r}))


if __name__ == "__main__":
    main()
