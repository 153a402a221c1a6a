#!/usr/bin/env python
"""Does the full-duplex host <-> device copy rate depend on where the pinned buffers land?

Prints the box's NUMA layout, then allocates pinned buffer pairs repeatedly -- unbound, and with the process bound to the
CPUs of each NUMA node before allocation and first touch -- and times one step's worth of bytes (16.8 MB) both ways at once.
"""
import glob
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def duplex_ms(torch, hin, hout, din, dout, s1, s2, reps=30):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        with torch.cuda.stream(s1):
            for _ in range(reps):
                din.copy_(hin, non_blocking=True)
            e1.record()
        with torch.cuda.stream(s2):
            for _ in range(reps):
                hout.copy_(dout, non_blocking=True)
            e2.record()
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e1), e0.elapsed_time(e2)) / reps)
    return best


def main():
    import torch
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        nodes[int(d.rsplit("node", 1)[1])] = cpulist(open(d + "/cpulist").read())
    print(json.dumps({"numa_nodes": {k: "%d cpus" % len(v) for k, v in nodes.items()}, "cpus": os.cpu_count()}))
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        print("gpu0 cpu affinity words:", [hex(w) for w in words])
    except Exception as ex:
        print("nvml:", ex)
    nbytes = 16 * 1024 * 1024
    din = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    all_cpus = sorted(os.sched_getaffinity(0))
    keep = []
    for label, cpus in [("unbound", all_cpus)] + [("node%d" % k, v) for k, v in nodes.items()]:
        cpus = [c for c in cpus if c in all_cpus]
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        res = []
        for _ in range(6):
            hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            hin.fill_(1)
            hout.fill_(2)
            keep.append((hin, hout))  # keep them so that every round gets fresh pages
            res.append(round(duplex_ms(torch, hin, hout, din, dout, s1, s2), 4))
        print(label, "duplex ms per 16.8 MB each way:", res)
    os.sched_setaffinity(0, all_cpus)
    # placement or time?  the same buffers again, three rounds
    for r in range(3):
        print("round %d, same buffers:" % r, [round(duplex_ms(torch, a, b, din, dout, s1, s2, reps=10), 3) for a, b in keep])
    # one direction at a time, per buffer
    def one_way(src, dst, reps=10):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        return round(best, 3)
    import bench
    a, b = keep[-1]
    a.fill_(3)
    b.fill_(4)
    print("last pair refilled: h2d %.3f d2h %.3f duplex %.3f" % (one_way(a, din), one_way(dout, b), duplex_ms(torch, a, b, din, dout, s1, s2, reps=10)))
    import time
    t0 = time.time()
    bench.evict_cpu_caches()
    print("evict_cpu_caches: %.2f s" % (time.time() - t0))
    print("after eviction:     h2d %.3f d2h %.3f duplex %.3f" % (one_way(a, din), one_way(dout, b), duplex_ms(torch, a, b, din, dout, s1, s2, reps=10)))
    print("h2d alone:", [one_way(a, din) for a, _ in keep])
    print("d2h alone:", [one_way(dout, b) for _, b in keep])


if __name__ == "__main__":
    main()
