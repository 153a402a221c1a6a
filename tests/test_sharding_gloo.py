"""N > 1 host logic on CPU: torch.distributed with the gloo backend, world_size 2.

The path shards by channel with no data-path collective.  Each rank takes its block-contiguous
shard from bbx_shard_range, runs the convolver on it (the CPU oracle stands in for the device here), and the
gathered outputs must be bit-identical to the unsharded run; the bench's barrier + max-over-ranks reduction is
exercised the same way bench.py uses it."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nch, ret):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bbcat_dsp_b200 as bbx
    import cpulibs as cl
    from convkit import OracleDriver, interleave, make_ir, make_noise, run_float

    B, L, nblk = 64, 200, 6
    first, count = bbx.shard_range(nch, rank, world)
    drv = OracleDriver(B, 4, count, max_blocks=3, max_delay=16, fractional_delay=True)
    for c in range(count):
        drv.select(c, drv.filter(make_ir(2000 + first + c, L)), delay=1.5 * (first + c))
    x = interleave([make_noise(1000 + first + c, nblk * B) for c in range(count)])
    y = run_float(drv, x, 3 * B)  # [frames][count]

    # gather the shards (sizes differ when nch % world != 0: pad to the largest shard)
    maxc = -(-nch // world)
    pad = np.zeros((nblk * B, maxc), dtype=np.float32)
    pad[:, :count] = y
    bufs = [torch.zeros(pad.shape, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(pad))
    # the bench's timing reduction: barrier, then MAX over ranks
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        parts = []
        for r in range(world):
            f, c = bbx.shard_range(nch, r, world)
            parts.append(bufs[r].numpy()[:, :c])
        ret["y"] = np.concatenate(parts, axis=1)
        ret["tmax"] = float(t.item())
    dist.destroy_process_group()


def _run(world, nch):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), nch, ret), nprocs=world, join=True)
    return ret["y"], ret["tmax"]


def test_channel_sharding_world2_bit_identical():
    sys.path.insert(0, HERE)
    from convkit import OracleDriver, interleave, make_ir, make_noise, run_float

    nch, B, L, nblk = 5, 64, 200, 6  # 5 channels over 2 ranks: shards of 3 and 2
    y_sharded, tmax = _run(2, nch)
    assert tmax == 2.0
    drv = OracleDriver(B, 4, nch, max_blocks=3, max_delay=16, fractional_delay=True)
    for c in range(nch):
        drv.select(c, drv.filter(make_ir(2000 + c, L)), delay=1.5 * c)
    x = interleave([make_noise(1000 + c, nblk * B) for c in range(nch)])
    y_full = run_float(drv, x, 3 * B)
    assert y_sharded.shape == y_full.shape
    assert np.array_equal(y_sharded.view(np.uint32), y_full.view(np.uint32))
