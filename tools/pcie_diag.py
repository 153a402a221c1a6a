#!/usr/bin/env python
"""Host<->device copy bandwidth with 1..N GPUs copying at the same time (torchrun, one rank per GPU).

Diagnostic for the e2e number of bench.py at N > 1: every rank moves the C3 step's PCM (16.8 MB each way) between
pinned host memory and its GPU on two streams, while only the ranks of the phase's active set are copying.  Shows
whether the host side (PCIe switch uplinks, IOMMU, NUMA placement of the pinned buffers) sustains N concurrent
full-duplex streams.  Rank 0 prints one JSON line per phase and the topology the box reports.
"""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    bound = None
    if os.environ.get("DIAG_BIND", "0") == "1":
        from bench import bind_near_gpu
        bound = bind_near_gpu(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    nbytes = 64 * 512 * 128 * 4
    hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hin.fill_(1)
    hout.fill_(2)
    din = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dout = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    if rank == 0:
        for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["numactl", "-H"]):
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
                print("#### " + " ".join(cmd) + "\n" + out, flush=True)
            except Exception as ex:
                print("#### %s failed: %s" % (cmd, ex), flush=True)
        print("#### affinity", sorted(os.sched_getaffinity(0)), "bound", bound, flush=True)
    phases = [[0]]
    if world >= 2:
        phases += [[0, 1]]
    if world >= 4:
        phases += [[0, 1, 2, 3]]
    if world >= 8:
        phases += [[0, 4], [0, 2, 4, 6], list(range(8))]
    for mode in ("h2d", "d2h", "both"):
        for act in phases:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            iters = 30
            gbs = 0.0
            if rank in act:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s1.wait_event(e0)
                s2.wait_event(e0)
                for _ in range(iters):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s1):
                            din.copy_(hin, non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s2):
                            hout.copy_(dout, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s1)
                torch.cuda.current_stream().wait_stream(s2)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                gbs = nbytes * iters / (ms * 1e-3) / 1e9  # per direction
            t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            if rank == 0:
                print(json.dumps({"mode": mode, "active": act, "bind": bound is not None,
                                  "GBps_per_direction": [round(float(v.item()), 1) for v in allv]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
