#!/usr/bin/env python
"""bench.py -- headline benchmark of the partitioned-convolution path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric / north_star targets are quoted on):
128-channel long reverb, 3 s IR (144000 taps) at 48 kHz, 512-sample partitions (P = 282), distinct IR per
channel, uniform white-noise input.  One "step" = one pass of the hot path over one batch of T = 64 blocks
(32768 frames, 0.683 s of audio) for every channel of the rank.  Per step the working set (148 MB of filter
spectra + 181 MB of FDL + 147 MB of partial sums) is larger than the 126 MB L2, so no L2 flush is needed
between iterations.

  value  channel-seconds of audio per second, inputs resident in HBM, K steps timed with CUDA events on the
         engine streams between barriers, max over ranks.  The engine's default path: calls with >= 8 blocks run
         the time-batched FDL MAC (k_fdl_mac_tb, FP32-bound), shorter calls the streaming MAC (k_fdl_mac, HBM-bound).
  e2e    the same metric through the C-ABI call bbx_process_async() with HOST (pinned) buffers: the H2D copy of
         the step's input and the D2H copy of its output are inside the timed region (copy streams overlap them
         with the kernels of the neighbouring steps)
  roofline            the dominant kernel of the timed region (k_fdl_mac_tb): FP32 FMA rate against the SIMT peak
  roofline_streaming  the streaming MAC (k_fdl_mac) timed in the same run: algorithmic HBM bytes against the
                      measured HBM peak -- the roofline BASELINE.json's north_star names
  roofline_mimo       the tensor-core kernel (k_mimo_tc) on BASELINE.json's MIMO config (C5), same run: TF32 MMA
                      rate against the measured tensor peak
N > 1    weak scaling: every rank runs its own 128-channel shard (rank r = channels 128r .. 128r+127 of a
         128N-channel renderer); channels are independent, so there is no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FS = 48000
B = 512
L = 144000
P = (L + B - 1) // B  # 282
NCH = 128
T = 64
K_BINS = B + 1
# algorithmic bytes of one channel-block-step of the FDL MAC (fp32 in / fp32 out)
BYTES_PER_CHANNEL_BLOCK = 16 * P * K_BINS + 16 * K_BINS + (4 + 4) * B  # 2,326,960
FLOPS_PER_CHANNEL_BLOCK = 8 * P * K_BINS  # 1,157,328 (complex MAC = 4 FMA = 8 flop)
WORKLOAD = "C3: 128-channel long reverb, 144000-tap IR (3 s @ 48 kHz), 512-sample partitions (P=282), f32 in/out"


def make_ir(seed, n):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal(n) * np.exp(-6.9 * np.arange(n) / n)
    h /= np.sqrt((h ** 2).sum())
    return h.astype(np.float32)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, n = [], None, set(), 0
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx, util = float(f[1]), float(f[2]), float(f[4])
            except ValueError:
                continue
            n += 1
            smax = mx
            if util >= 50:
                sm.append(clk)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": n, "samples_under_load": len(sm)}


def cpu_convolver_rate(n_blocks, nthreads, want_seconds=None):
    """Time the CPU oracle convolver (fp32 UPOLS, SURVEY.md 8.A, one worker per channel up to nthreads) on the
    C3 workload for n_blocks block-steps.  Returns (channel-s/s, seconds, blocks)."""
    import cpulibs
    orc = cpulibs.oracle()
    cv = orc.convolver(block=B, max_partitions=P, n_inputs=NCH, ring_len=4 * B, nthreads=nthreads)
    filters = []
    for c in range(NCH):
        f = orc.filter(make_ir(2000 + c, L), B)
        filters.append(f)
        cv.set_filter(c, f, False, 0.0)
    rng = np.random.default_rng(1000)
    x = rng.uniform(-1, 1, (4 * B, NCH)).astype(np.float32)
    cv.process(x, cpulibs.FMT_FLOAT, NCH, cpulibs.FMT_FLOAT, NCH, 4 * B)  # warm-up: page in spectra
    done, t0 = 0, time.perf_counter()
    while True:
        cv.process(x, cpulibs.FMT_FLOAT, NCH, cpulibs.FMT_FLOAT, NCH, 4 * B)
        done += 4
        el = time.perf_counter() - t0
        if done >= n_blocks and (not want_seconds or el >= want_seconds):
            break
    return NCH * done * B / FS / el, el, done


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def reference_function_rates():
    """the in-tree reference functions of the path timed single-threaded on this host -- the reference's
    own code when oracle/_ref was built, else the C restatement.  Sanity anchors next to the convolver baseline."""
    import cpulibs
    lib, kind = cpulibs.reference(), "reference"
    if lib is None:
        lib, kind = cpulibs.oracle(), "port"
    nch, nfr = 32, 48000
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, nch * nfr).astype(np.float32)
    s24 = np.zeros(nch * nfr * 3, dtype=np.uint8)
    back = np.zeros(nch * nfr, dtype=np.float32)

    def rate(fn, samples):
        fn()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 0.4:
            fn()
            n += 1
        return samples * n / (time.perf_counter() - t0) / 1e6

    out = {"kind": kind, "threads": 1, "unit": "Msample/s"}
    out["transfer_f32_to_s24_32ch"] = rate(lambda: lib.transfer(x.view(np.uint8), cpulibs.FMT_FLOAT, 0, 0, nch, s24, cpulibs.FMT_24, 0, 0,
                                                                nch, nch, nfr), nch * nfr)
    out["transfer_s24_to_f32_32ch"] = rate(lambda: lib.transfer(s24, cpulibs.FMT_24, 0, 0, nch, back.view(np.uint8), cpulibs.FMT_FLOAT, 0,
                                                                0, nch, nch, nfr), nch * nfr)
    bus = np.zeros(2 * nfr, dtype=np.float32)
    mono = x[:nfr].copy()

    def mix32():
        for p in range(32):
            lib.mix(mono, 0, 1, bus, p & 1, 2, 1, nfr, 0.5)
    out["mix_32_mono_paths_to_stereo"] = rate(mix32, 32 * nfr)
    ring = x[:4096].copy()
    pos = rng.uniform(16, 4000, 20000)
    out["fractional_sample_f32"] = rate(lambda: lib.frac(ring, 0, 1, 4096, pos), pos.size)
    return out


def mimo_leg(bbx, torch, device, steps=100):
    """BASELINE.json configs[4] (C5): 64-in x 64-out matrix of 4096-tap IRs, B = 512, 64-block steps.  The per-bin
    complex GEMM runs on the tensor cores (k_mimo_tc: tcgen05.mma kind::tf32, 3 MMAs per product for fp32 accuracy).
    achieved = TF32 MMA flops issued per launch / the kernel's average launch time (CUDA events around every launch);
    peak = the measured dense bf16 rate of MEASURED_PEAKS.json / 2 (TF32 runs at half the bf16 rate)."""
    nin = nout = 64
    Lm, Pm, Tm = 4096, 8, 64
    eng = bbx.Convolver(B, Pm, nin, n_outputs=nout, mode=bbx.MODE_MIMO, max_blocks=Tm, device=device)
    for o in range(nout):
        for i in range(nin):
            eng.SelectFilter(o * nin + i, eng.CreateFilter(make_ir(2000 + 64 * o + i, Lm)))
    frames = Tm * B
    x = (torch.rand((frames, nin), device="cuda") * 2 - 1).contiguous()
    y = torch.empty((frames, nout), device="cuda", dtype=torch.float32)
    for _ in range(5):
        eng.ConvolveDev(x.data_ptr(), bbx.FMT_FLOAT, nin, y.data_ptr(), bbx.FMT_FLOAT, nout, frames)
    eng.Sync()
    eng.profile_mac(True)
    eng.timer_start()
    for _ in range(steps):
        eng.ConvolveDev(x.data_ptr(), bbx.FMT_FLOAT, nin, y.data_ptr(), bbx.FMT_FLOAT, nout, frames)
    ms = eng.timer_stop()
    mac = eng.mac_time()
    eng.profile_mac(False)
    n_tc, status = eng.tensor_status()
    eng.close()
    lms = mac["ms"] / max(1, mac["launches"])
    # per bin: M = 128 (64 outputs x re/im), K = 2 * nin * P = 1024, N = 64 block-steps, 3 TF32 MMAs per product
    flops = 3 * 2 * 128 * (2 * nin * Pm) * Tm * B
    peak_bf16, src = None, None
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            peak_bf16, src = float(json.load(open(path))["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 = half the bf16 rate)"
        except Exception:
            pass
    if peak_bf16 is None:
        peak_bf16, src = 2250.0, "nominal dense bf16 2.25 PF / 2 (TF32)"
    peak = peak_bf16 / 2
    ach = flops / (lms * 1e-3) / 1e12 if lms > 0 else 0.0
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": 180.8e6,
            "kernel": "k_mimo_tc<6>", "launch_ms": lms, "tf32_flops_per_launch": flops, "peak_source": src,
            "workload": "C5: MIMO 64 in x 64 out, 4096-tap matrix, B=512, 64-block steps, f32 in/out",
            "value": nout * steps * frames / FS / (ms * 1e-3), "unit_value": "output-channel-s/s", "ms_per_step": ms / steps,
            "tensor_launches": n_tc, "status": status, "mac_share_of_step": mac["ms"] / ms if ms > 0 else None,
            "useful_fp32_equivalent_TFLOPs": flops / 3 / (lms * 1e-3) / 1e12 if lms > 0 else 0.0,
            "note": "traffic = dram__bytes_read + write of one launch (ncu --set full, profiles/); 134 MB of spectra are read "
                    "once per 64 block-steps"}


def bind_near_gpu(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU before any pinned host memory is allocated
    (first touch then places the staging buffers on that NUMA node).  Returns the CPU list or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  BlockConvolver/Convolver are absent
    from the mounted bbcat-dsp tree and FFTW is not installed, so this is the oracle port
    (oracle/upols.c + convolver.c, OpenMP over channels) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    blocks_per_step = 8  # bounded sample of the 64-block step
    import cpulibs
    orc = cpulibs.oracle()
    cv = orc.convolver(block=B, max_partitions=P, n_inputs=NCH, ring_len=4 * B, nthreads=cores)
    keep = []
    for c in range(NCH):
        f = orc.filter(make_ir(2000 + c, L), B)
        keep.append(f)
        cv.set_filter(c, f, False, 0.0)
    x = np.random.default_rng(1000).uniform(-1, 1, (blocks_per_step * B, NCH)).astype(np.float32)
    for _ in range(args.warmup):
        cv.process(x, cpulibs.FMT_FLOAT, NCH, cpulibs.FMT_FLOAT, NCH, blocks_per_step * B)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cv.process(x, cpulibs.FMT_FLOAT, NCH, cpulibs.FMT_FLOAT, NCH, blocks_per_step * B)
    el = time.perf_counter() - t0
    value = NCH * args.steps * blocks_per_step * B / FS / el
    sample = "%d steps x %d blocks x %d channels of the C3 workload (step bounded from %d to %d blocks)" % (
        args.steps, blocks_per_step, NCH, T, blocks_per_step)
    line = {
        "impl": "reference", "metric": "channel_seconds_per_second", "value": value, "unit": "channel-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "channels": NCH, "block": B, "partitions": P, "blocks_per_step": blocks_per_step,
                   "note": "CPU port of the absent BlockConvolver/Convolver (own FFT, FFTW unavailable)"},
        "cpu_baseline": {"value": value, "unit": "channel-s/s", "cores": cores, "kind": "port", "nproc": os.cpu_count(),
                         "cpu_model": cpu_model(), "sample": sample},
        "e2e": {"value": value, "unit": "channel-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--occ", type=int, default=0, help="MAC CTAs per SM (tuning)")
    ap.add_argument("--l2keep", type=int, default=0, help="sixteenths of H/FDL lines kept L2-resident (streaming MAC), 0 = default")
    ap.add_argument("--tile", type=int, default=0, help="time-batched MAC tile: 0 = default (16), 1 = streaming kernel only, 16, 32")
    ap.add_argument("--no-streaming", action="store_true", help="skip the streaming-MAC roofline pass")
    ap.add_argument("--blocks", type=int, default=T, help="blocks per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-mimo", action="store_true", help="skip the C5 MIMO tensor-core leg (roofline_mimo)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import bbcat_dsp_b200 as bbx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ndev = bbx.device_count()
    if ndev < 1:
        raise SystemExit("bench.py needs a CUDA device: libbbx has no CPU fallback")
    # Rank -> GPU.  With more visible GPUs than ranks the ranks are spread evenly over the box (N = 2 on an 8-GPU box: GPUs 0
    # and 4): the path has no GPU-to-GPU traffic, and the host-to-device copies of the e2e leg then share fewer PCIe switch
    # uplinks (profiles/r01_pcie_diag_n8.txt: two GPUs copying both ways get 26 GB/s each on GPUs 0,1 and 34 / 42 on 0,4).
    stride = ndev // world if (world > 1 and ndev >= world) else 1
    local = (local_rank * stride) % ndev
    torch.cuda.set_device(local)
    numa = bind_near_gpu(local) if world > 1 else None  # pinned staging on the GPU's own NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    nblk = args.blocks
    warmup = max(3, args.warmup)

    # ---- set up the rank's shard: 128 channels, distinct IRs, noise resident in HBM ----
    eng = bbx.Convolver(B, P, NCH, max_blocks=nblk, device=local, mac_ctas_per_sm=args.occ,
                        mac_l2_keep_16ths=args.l2keep, mac_time_tile=args.tile)
    for c in range(NCH):
        eng.SelectFilter(c, eng.CreateFilter(make_ir(2000 + NCH * rank + c, L)))
    frames = nblk * B
    g = torch.Generator(device="cuda")
    g.manual_seed(1000 + rank)
    x_dev = (torch.rand((frames, NCH), device="cuda", generator=g) * 2 - 1).contiguous()
    y_dev = torch.empty((frames, NCH), device="cuda", dtype=torch.float32)
    in_bytes = frames * NCH * 4
    # two pinned buffer pairs: the asynchronous host API overlaps the copies of one step with the kernels of the next
    hins = [bbx.PinnedBuffer(in_bytes) for _ in range(2)]
    houts = [bbx.PinnedBuffer(in_bytes) for _ in range(2)]
    for i, h in enumerate(hins):
        h.array[:] = np.random.default_rng(1000 + rank + 17 * i).uniform(-1, 1, frames * NCH).astype(np.float32).view(np.uint8)
    hin, hout = hins[0], houts[0]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_dev():
        eng.ConvolveDev(x_dev.data_ptr(), bbx.FMT_FLOAT, NCH, y_dev.data_ptr(), bbx.FMT_FLOAT, NCH, frames)

    def step_host(i):
        eng.ConvolveHostPtrAsync(hins[i & 1].ptr, bbx.FMT_FLOAT, NCH, houts[i & 1].ptr, bbx.FMT_FLOAT, NCH, frames)

    sampler = ClockSampler(local)
    for _ in range(warmup):
        step_dev()
    eng.Sync()
    if rank == 0:
        sampler.start()

    # ---- value: device-resident inputs ----
    barrier()
    eng.profile_mac(True)
    l0 = eng.launch_count()
    eng.timer_start()
    for _ in range(args.steps):
        step_dev()
    ms = eng.timer_stop()
    barrier()
    launches = eng.launch_count() - l0
    mac = eng.mac_time()
    eng.profile_mac(False)
    ms = max_over_ranks(ms)
    audio_s = NCH * frames / FS  # channel-seconds per step per rank
    value = world * audio_s * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through bbx_process ----
    for i in range(4):
        step_host(i)
    eng.Sync()
    barrier()
    eng.timer_start()
    for i in range(args.steps):
        step_host(i)
    ms_e2e = eng.timer_stop()  # covers the last D2H copy
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    e2e = world * audio_s * args.steps / (ms_e2e * 1e-3)
    clocks = None
    if rank == 0:
        if len(sampler.lines) < 8:
            # the timed loops were shorter than a few nvidia-smi periods (small --steps): keep the identical load
            # running for ~1.5 s more so that the clock record has samples under load; noted in the record
            t_end = time.perf_counter() + 1.5
            while time.perf_counter() < t_end:
                for _ in range(50):
                    step_dev()
                eng.Sync()
            clocks = sampler.stop()
            clocks["window"] = "timed loops + 1.5 s of identical steps (timed region shorter than the sampling period)"
        else:
            clocks = sampler.stop()
            clocks["window"] = "value and e2e timed loops"

    # ---- per-block latency, streaming T = 1 through the host API ----
    latency = None
    if not args.no_latency and rank == 0:
        def block_latency(nlat=1000, warm=100):  # 1000+ steps after 100 warm-up
            lat = []
            for i in range(nlat + warm):
                t0 = time.perf_counter()
                eng.ConvolveHostPtr(hin.ptr, bbx.FMT_FLOAT, NCH, hout.ptr, bbx.FMT_FLOAT, NCH, B)
                if i >= warm:
                    lat.append(time.perf_counter() - t0)
            lat = np.array(lat) * 1e6
            return float(np.percentile(lat, 50)), float(np.percentile(lat, 99)), nlat
        p50, p99, nlat = block_latency()
        latency = {"p50_us": p50, "p99_us": p99, "blocks": nlat, "block_period_us": 1e6 * B / FS,
                   "mode": "T=1, bbx_process with pinned host buffers (direct path: the PCM kernels read / write the "
                           "pinned buffers over PCIe, no copy-engine hops), host clock",
                   "direct_calls": eng.direct_calls()}
        eng.set_direct_io(0)  # the same call through the staged copy-engine pipeline, for comparison
        p50, p99, _ = block_latency()
        eng.set_direct_io(1 << 20)
        latency["staged_p50_us"], latency["staged_p99_us"] = p50, p99

    # ---- roofline of the dominant kernel of the timed region, CUDA events around every MAC launch ----
    peak, peak_src = peaks()
    units_per_launch = mac["channel_blocks"] / max(1, mac["launches"])
    mac_ms = mac["ms"] / max(1, mac["launches"])
    traffic_tb = traffic_stream = None
    tpath = os.path.join(ROOT, "profiles", "mac_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic_stream = tj["dram_bytes_per_channel_block"]
            traffic_tb = tj.get("tb_dram_bytes_per_channel_block")
        except Exception:
            pass
    batched = args.tile != 1
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    fp32_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12  # TFLOP/s: 148 SMs x 128 FMA lanes x 2 flop x max SM clock
    # the FP32 ceiling this GPU actually reaches (MEASURED_PEAKS.json has no FP32 figure): a pure packed-FMA kernel with
    # the MAC's operand pattern, measured in this run -- burst (best isolated launch) and sustained (0.5 s back to back)
    fp32_burst, fp32_sustained = bbx.probe_fp32_tflops(local, 0.5)
    fp32_peak = fp32_burst if fp32_burst > 0 else fp32_nominal
    if batched:
        tflops = FLOPS_PER_CHANNEL_BLOCK * units_per_launch / (mac_ms * 1e-3) / 1e12 if mac_ms > 0 else 0.0
        roofline = {"bound": "fp32", "achieved": tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tflops / fp32_peak,
                    "traffic": traffic_tb * units_per_launch if traffic_tb else None,
                    "kernel": "k_fdl_mac_tb<16,256,8>" if args.tile in (0, 16) else "k_fdl_mac_tb<32,256,8>",
                    "launch_ms": mac_ms, "units_per_launch": units_per_launch,
                    "flops_per_launch": FLOPS_PER_CHANNEL_BLOCK * units_per_launch,
                    "peak_source": "measured in this run: pure packed-FMA probe kernel (bbx_probe_fp32_tflops), burst = best of 5 "
                                   "isolated launches (MEASURED_PEAKS.json has no FP32 figure)",
                    "peak_sustained": fp32_sustained, "frac_sustained": tflops / fp32_sustained if fp32_sustained > 0 else None,
                    "peak_nominal": fp32_nominal, "frac_nominal": tflops / fp32_nominal,
                    "peak_nominal_source": "148 SM x 128 lanes x 2 flop x %.0f MHz" % sm_max,
                    "note": "time-batched MAC: H[p] is loaded once per 16 block-steps, 4 FMA per loaded byte -> bound by the "
                            "FP32 pipe, not HBM; the hbm roofline of the streaming kernel is in roofline_streaming",
                    "hbm_equivalent_GBps": BYTES_PER_CHANNEL_BLOCK * units_per_launch / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0,
                    "mac_share_of_step": mac["ms"] / ms if ms > 0 else None}
    else:
        achieved = BYTES_PER_CHANNEL_BLOCK * units_per_launch / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic_stream * units_per_launch if traffic_stream else None, "kernel": "k_fdl_mac",
                    "launch_ms": mac_ms, "units_per_launch": units_per_launch,
                    "algorithmic_bytes_per_launch": BYTES_PER_CHANNEL_BLOCK * units_per_launch, "peak_source": peak_src,
                    "mac_share_of_step": mac["ms"] / ms if ms > 0 else None}

    # ---- the streaming MAC in the same run (the HBM-bound kernel the north_star roofline is about) ----
    roofline_streaming = None
    if batched and not args.no_streaming:
        eng.set_tuning(time_tile=1)
        ks = max(3, min(args.steps, 40))
        for _ in range(3):
            step_dev()
        barrier()
        eng.profile_mac(True)
        eng.timer_start()
        for _ in range(ks):
            step_dev()
        ms_s = max_over_ranks(eng.timer_stop())
        barrier()
        mac_s = eng.mac_time()
        eng.profile_mac(False)
        eng.set_tuning(time_tile=args.tile or 16)
        upl = mac_s["channel_blocks"] / max(1, mac_s["launches"])
        lms = mac_s["ms"] / max(1, mac_s["launches"])
        ach = BYTES_PER_CHANNEL_BLOCK * upl / (lms * 1e-3) / 1e9 if lms > 0 else 0.0
        roofline_streaming = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                              "traffic": traffic_stream * upl if traffic_stream else None, "kernel": "k_fdl_mac",
                              "launch_ms": lms, "units_per_launch": upl,
                              "algorithmic_bytes_per_launch": BYTES_PER_CHANNEL_BLOCK * upl, "peak_source": peak_src,
                              "value_streaming": world * audio_s * ks / (ms_s * 1e-3), "steps": ks,
                              "mac_share_of_step": mac_s["ms"] / ms_s if ms_s > 0 else None}

    # ---- the tensor-core kernel of the path (C5 MIMO, k_mimo_tc) in the same run: rank 0, N = 1 ----
    roofline_mimo = None
    if rank == 0 and world == 1 and not args.no_mimo:
        roofline_mimo = mimo_leg(bbx, torch, local)

    # ---- CPU baseline on rank 0 at N = 1 (bounded sample of the same workload) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        rate, secs, blocks = cpu_convolver_rate(64, cores, want_seconds=12.0)
        cpu = {"value": rate, "unit": "channel-s/s", "cores": cores, "kind": "port", "nproc": os.cpu_count(), "cpu_model": cpu_model(),
               "reference_functions": reference_function_rates(),
               "sample": "%d block-steps x %d channels of the C3 workload in %.1f s (oracle UPOLS, OpenMP over channels, own FFT)" % (
                   blocks, NCH, secs)}

    if rank == 0:
        line = {
            "metric": "channel_seconds_per_second", "value": value, "unit": "channel-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "channels_per_gpu": NCH, "total_channels": NCH * world, "block": B,
                       "partitions": P, "blocks_per_step": nblk,
                       "mac": "time-batched (tile 16)" if args.tile in (0, 16) else ("streaming" if args.tile == 1 else "time-batched (tile %d)" % args.tile),
                       "l2": "inputs larger than L2: 148 MB spectra + 181 MB FDL + partial sums per step vs 126 MB L2, no flush",
                       "parallelism": "channel-sharded x%d, no collective" % world,
                       "devices": "rank r on GPU %d*r of %d visible" % (stride, ndev),
                       "host_binding": ("rank 0 bound to %d CPUs local to its GPU" % len(numa)) if numa else "none"},
            "x_realtime_per_channel": value / (NCH * world),
            "roofline": roofline, "roofline_streaming": roofline_streaming, "roofline_mimo": roofline_mimo, "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "channel-s/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": in_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "latency": latency,
        }
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
