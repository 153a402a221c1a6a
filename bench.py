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

  value   channel-seconds of audio per second, inputs resident in HBM, K steps timed with CUDA events on the
          engine streams between barriers, max over ranks.  The engine's default path: calls with >= 8 blocks run
          the time-batched FDL MAC (FP32-bound), shorter calls the streaming MAC (k_fdl_mac, HBM-bound).
  e2e     the same metric through the C-ABI call bbx_process_async() with HOST (pinned) buffers: the H2D copy of
          the step's input and the D2H copy of its output are inside the timed region (copy streams overlap them
          with the kernels of the neighbouring steps; the synthetic host buffers are pushed out of the CPU caches first,
          see evict_cpu_caches).  e2e.roofline = the PCIe bytes of a step against the
          host<->device copy rate measured in the same run with every rank copying (the bound of this leg).
  parity  the timed configuration checked outside the timed region: a window of the engine's output against a
          float64 direct convolution (numpy) -- "parity_pin": "definition", because BlockConvolver / Convolver are
          absent from the reference tree, so there is no BBC output to pin against.
  roofline            the dominant kernel of the timed region (time-batched FDL MAC): FP32 FMA rate; frac is
                      against the NOMINAL FP32 peak (148 SM x 128 lanes x 2 x max SM clock), frac_probe against a
                      pure packed-FMA kernel measured in this run
  roofline_streaming  the streaming MAC (k_fdl_mac) timed in the same run, the HBM roofline BASELINE.json's
                      north_star names: frac = DRAM bytes (ncu capture under profiles/, scaled to this launch) per
                      second against the measured HBM peak; frac_algorithmic uses SURVEY.md 8(d)'s algorithmic bytes
  roofline_mimo       the tensor-core kernel (k_mimo_tc) on BASELINE.json's MIMO config (C5), same run
  configs             N = 1: BASELINE.json's five configs, each with throughput, T = 1 latency and snr_db
N > 1     weak scaling of C3 (the headline line: every rank runs its own 128-channel shard, no collective) plus
          strong_scaling (128 / N channels per rank) and the C5 MIMO legs of SURVEY.md 8(e): input-sharded with the
          peer-memory NVLink mixdown and with ncclReduceScatter, and output-sharded (no collective), each with an
          in-run SNR of rank 0's outputs against the float64 direct convolution.
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
PARITY_PIN = "definition"  # float64 direct convolution: BlockConvolver/Convolver/FFTW are absent from the reference tree


def c3_config(world, nblk):
    """the workload-defining keys, identical in both arms (b200 and reference)"""
    return {"workload": WORKLOAD, "channels_per_gpu": NCH, "total_channels": NCH * world, "block": B, "partitions": P,
            "blocks_per_step": nblk,
            "l2": "working set per step (148 MB spectra + 181 MB FDL + partial sums) larger than the 126 MB L2: no flush"}


def make_ir(seed, n):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal(n) * np.exp(-6.9 * np.arange(n) / n)
    h /= np.sqrt((h ** 2).sum())
    return h.astype(np.float32)


def make_noise(seed, n):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n).astype(np.float32)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md: 6.65 TB/s)", "bf16_tflops": 1590.0,
           "bf16_src": "fallback (B200_PROFILING.md: 1.59 PFLOP/s burst)"}
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            out.update(hbm_gbs=float(j["hbm_gbs"]), hbm_src="measured (MEASURED_PEAKS.json hbm_gbs)",
                       bf16_tflops=float(j["bf16_tflops"]), bf16_src="measured (MEASURED_PEAKS.json bf16_tflops, burst)")
        except Exception:
            pass
    return out


# ---- parity: float64 direct convolution of a window (numpy; the definition of the path) ----------------
def direct_window(terms, n0, n1):
    """y[n] = sum over terms of g * sum_j h[j] x[n - d - j] for n in [n0, n1), zero history; terms = (x, h, d, g)."""
    out = np.zeros(n1 - n0)
    for x, h, d, g in terms:
        h64 = np.asarray(h[:max(0, n1 - d)], dtype=np.float64)
        if h64.size == 0:
            continue
        lo, hi = n0 - d - (h64.size - 1), n1 - d
        seg = np.zeros(hi - lo)
        a, b = max(lo, 0), min(hi, len(x))
        if b > a:
            seg[a - lo:b - lo] = x[a:b]
        out += g * np.convolve(seg, h64, mode="valid")
    return out


def compare(y, ref):
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = y - ref
    peak = float(np.abs(ref).max()) or 1.0
    p_sig, p_err = float((ref ** 2).sum()), float((err ** 2).sum())
    snr = 200.0 if p_err == 0 else 10.0 * np.log10(p_sig / p_err)
    return {"snr_db": float(snr), "max_abs_over_peak": float(np.abs(err).max() / peak)}


def worst(results):
    return {"snr_db": min(r["snr_db"] for r in results), "max_abs_over_peak": max(r["max_abs_over_peak"] for r in results),
            "tolerance": "snr_db >= 110 and max_abs_over_peak <= 1e-5 (BASELINE.json north_star)",
            "ok": all(r["snr_db"] >= 110.0 and r["max_abs_over_peak"] <= 1e-5 for r in results)}


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


# ---- CPU arms ----------------------------------------------------------------------------------------------------------
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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  BlockConvolver/Convolver are absent
    from the mounted bbcat-dsp tree and FFTW is not installed, so this is the oracle port
    (oracle/upols.c + convolver.c, OpenMP over channels) on the box's host cores, on the same configuration as the
    b200 arm: 128 channels, 64-block steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = max(1, args.gpus)
    cores = host_cores()
    nblk = args.blocks
    import cpulibs
    orc = cpulibs.oracle()
    cv = orc.convolver(block=B, max_partitions=P, n_inputs=NCH, ring_len=(nblk + 1) * B, nthreads=cores)
    keep = []
    for c in range(NCH):
        f = orc.filter(make_ir(2000 + c, L), B)
        keep.append(f)
        cv.set_filter(c, f, False, 0.0)
    x = np.random.default_rng(1000).uniform(-1, 1, (nblk * B, NCH)).astype(np.float32)
    for _ in range(args.warmup):
        cv.process(x, cpulibs.FMT_FLOAT, NCH, cpulibs.FMT_FLOAT, NCH, nblk * B)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cv.process(x, cpulibs.FMT_FLOAT, NCH, cpulibs.FMT_FLOAT, NCH, nblk * B)
    el = time.perf_counter() - t0
    value = NCH * args.steps * nblk * B / FS / el
    sample = "%d steps x %d blocks x %d channels of the C3 workload (one rank's shard; the CPU rate does not depend on N)" % (
        args.steps, nblk, NCH)
    line = {
        "impl": "reference", "metric": "channel_seconds_per_second", "value": value, "unit": "channel-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": c3_config(world, nblk),
        "note": "CPU port of the absent BlockConvolver/Convolver (own FFT, FFTW unavailable), OpenMP over channels",
        "cpu_baseline": {"value": value, "unit": "channel-s/s", "cores": cores, "kind": "port", "nproc": os.cpu_count(),
                         "cpu_model": cpu_model(), "sample": sample},
        "e2e": {"value": value, "unit": "channel-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "parity_pin": PARITY_PIN,
    }
    print(json.dumps(line))
    return 0


# ---- shared GPU helpers ----------------------------------------------------------------------------------------------
def block_latency(eng, bbx, fmt_in, in_ch, fmt_out, out_ch, blk, n=10000, warm=200, seed=5):
    """T = 1 streaming call through bbx_process with pinned host buffers, host clock: p50 / p99 over n blocks."""
    hin = bbx.PinnedBuffer(blk * in_ch * bbx.FMT_BYTES[fmt_in])
    hout = bbx.PinnedBuffer(blk * out_ch * bbx.FMT_BYTES[fmt_out])
    rng = np.random.default_rng(seed)
    if fmt_in < 4:
        hin.array[:] = rng.integers(0, 255, hin.nbytes, dtype=np.uint8)
    else:
        hin.array[:] = rng.uniform(-1, 1, blk * in_ch).astype(np.float32).view(np.uint8)
    lat = np.empty(n)
    d0 = eng.direct_calls()
    for i in range(n + warm):
        t0 = time.perf_counter()
        eng.ConvolveHostPtr(hin.ptr, fmt_in, in_ch, hout.ptr, fmt_out, out_ch, blk)
        if i >= warm:
            lat[i - warm] = time.perf_counter() - t0
    lat *= 1e6
    # the same calls timed inside libbbx (bbx_block_latency: CLOCK_MONOTONIC around each bbx_process): the latency at the
    # C ABI, which is what a C++ host of the reference sees; the loop above adds a Python / ctypes call per block
    clat = eng.BlockLatency(hin.ptr, fmt_in, in_ch, hout.ptr, fmt_out, out_ch, blk, n, warm)
    r = {"p50_us": float(np.percentile(clat, 50)), "p99_us": float(np.percentile(clat, 99)), "max_us": float(clat.max()), "blocks": n,
         "python_p50_us": float(np.percentile(lat, 50)), "python_p99_us": float(np.percentile(lat, 99)),
         "block_period_us": 1e6 * blk / FS, "direct_calls": eng.direct_calls() - d0,
         "mode": "T=1, bbx_process with pinned host buffers, host clock around the synchronous call; p50 / p99 / max = timed at "
                 "the C ABI (bbx_block_latency), python_* = the same call through ctypes"}
    hin.close()
    hout.close()
    return r


def launch_sync_floor(torch, n=2000):
    """host round trip of the smallest possible GPU call on this box -- one trivial kernel on a stream + a stream
    synchronise -- the floor under every per-block latency below (virtualised hosts sit well above bare metal)"""
    x = torch.zeros(1, device="cuda")
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    lat = np.empty(n)
    for i in range(n + 200):
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            x.add_(1)
        s.synchronize()
        if i >= 200:
            lat[i - 200] = time.perf_counter() - t0
    lat *= 1e6
    return {"p50_us": float(np.percentile(lat, 50)), "min_us": float(lat.min()),
            "what": "torch: one 1-element kernel on a side stream + stream.synchronize(), host clock"}


def timed_steps(eng, step, steps, warm=5):
    for i in range(warm):
        step(i)
    eng.Sync()
    l0 = eng.launch_count()
    eng.timer_start()
    for i in range(steps):
        step(i + warm)
    ms = eng.timer_stop()
    return ms, eng.launch_count() - l0


def dev_bytes(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).cuda()


# ---- BASELINE.json configs C1, C2, C4, C5 on one GPU -----------------------------------------------------------------
def leg_c1(bbx, torch, dev, steps):
    Bc, Lc, nch, Tc = 1024, 8192, 2, 64
    eng = bbx.Convolver(Bc, Lc // Bc, nch, max_blocks=Tc, device=dev)
    irs = [make_ir(2000 + c, Lc) for c in range(nch)]
    for c in range(nch):
        eng.SelectFilter(c, eng.CreateFilter(irs[c]))
    frames = Tc * Bc
    x = np.stack([make_noise(1000 + c, frames) for c in range(nch)], axis=1)
    xd, yd = dev_bytes(torch, x), torch.empty(frames * nch * 4, dtype=torch.uint8, device="cuda")
    eng.ConvolveDev(xd.data_ptr(), 4, nch, yd.data_ptr(), 4, nch, frames)
    eng.Sync()
    y = yd.cpu().numpy().view(np.float32).reshape(frames, nch)
    n0, n1 = frames - 4096, frames
    par = worst([compare(y[n0:n1, c], direct_window([(x[:, c], irs[c], 0, 1.0)], n0, n1)) for c in range(nch)])
    ms, launches = timed_steps(eng, lambda i: eng.ConvolveDev(xd.data_ptr(), 4, nch, yd.data_ptr(), 4, nch, frames), steps)
    r = {"config": "C1: stereo 2ch 48 kHz float, 8192-tap IR, 1024-sample partitions", "channels": nch,
         "value": nch * steps * frames / FS / (ms * 1e-3), "unit": "channel-s/s", "ms_per_step": ms / steps, "blocks_per_step": Tc,
         "launches_per_step": launches / steps, "parity": par, "snr_db": par["snr_db"],
         "latency": block_latency(eng, bbx, 4, nch, 4, nch, Bc)}
    eng.close()
    return r


def leg_c2(bbx, torch, dev, steps):
    Bc, Lc, nsrc, Tc = 256, 512, 64, 64
    eng = bbx.Convolver(Bc, Lc // Bc, nsrc, n_outputs=2, n_paths=2 * nsrc, mode=bbx.MODE_ROUTED, max_blocks=Tc, max_delay=48, device=dev)
    irs, delays, gain = {}, {}, 1.0 / 8
    for s in range(nsrc):
        for ear in range(2):
            p = 2 * s + ear
            irs[p], delays[p] = make_ir(2000 + p, Lc), (s * (1 + ear)) % 40
            eng.SetRoute(p, s, ear, gain)
            eng.SelectFilter(p, eng.CreateFilter(irs[p]), delay=float(delays[p]))
    frames = Tc * Bc
    x = np.stack([make_noise(1000 + s, frames) for s in range(nsrc)], axis=1)
    xd, yd = dev_bytes(torch, x), torch.empty(frames * 2 * 4, dtype=torch.uint8, device="cuda")
    eng.ConvolveDev(xd.data_ptr(), 4, nsrc, yd.data_ptr(), 4, 2, frames)
    eng.Sync()
    y = yd.cpu().numpy().view(np.float32).reshape(frames, 2)
    n0, n1 = frames - 2048, frames
    par = worst([compare(y[n0:n1, ear], direct_window([(x[:, s], irs[2 * s + ear], delays[2 * s + ear], gain) for s in range(nsrc)], n0, n1))
                 for ear in range(2)])
    ms, launches = timed_steps(eng, lambda i: eng.ConvolveDev(xd.data_ptr(), 4, nsrc, yd.data_ptr(), 4, 2, frames), steps)
    r = {"config": "C2: 64-source binaural renderer, 64 inputs x 2 ears (128 paths), 512-tap HRIRs, 256-sample blocks, ITD delays, "
                   "time-domain mixdown", "channels": nsrc, "paths": 2 * nsrc,
         "value": nsrc * steps * frames / FS / (ms * 1e-3), "unit": "source-s/s", "ms_per_step": ms / steps, "blocks_per_step": Tc,
         "launches_per_step": launches / steps, "parity": par, "snr_db": par["snr_db"],
         "latency": block_latency(eng, bbx, 4, nsrc, 4, 2, Bc)}
    eng.close()
    return r


def c4_schedule(m, nch, nbank):
    """at switch m channel c selects IR (m + c) mod 16 with delay 16 + 37.3 ((7 m + c) mod 11) / 11; the call
    that follows covers the blocks up to the next switch (every 100 ms = block index ceil(4800 m / 512))"""
    sel = [(m + c) % nbank for c in range(nch)]
    dly = [16 + 37.3 * ((m * 7 + c) % 11) / 11 for c in range(nch)]
    cd = lambda a: -(-a // 512)
    return sel, dly, cd((m + 1) * 4800) - cd(m * 4800)


def leg_c4(bbx, torch, dev, steps):
    """32 channels x bank of 16 IRs, IR select every 100 ms with crossfade + fractional delay, s24 in/out.  Parity of the
    full-size configuration against the CPU oracle convolver (the checker; switching + fractional delays have no closed
    float64 form), outputs compared as floats plus the <= 1 LSB(24) rule of SURVEY.md 8.A."""
    import cpulibs
    from parity import s24_to_float
    Bc, Lc, nch, nbank, Tmax = 512, 4096, 32, 16, 10
    eng = bbx.Convolver(Bc, Lc // Bc, nch, max_blocks=Tmax, max_delay=64, fractional_delay=True, device=dev)
    irs = [[make_ir(2000 + 16 * c + k, Lc) for k in range(nbank)] for c in range(nch)]
    bank = [[eng.CreateFilter(irs[c][k]) for k in range(nbank)] for c in range(nch)]
    orc = cpulibs.oracle()
    ocv = orc.convolver(block=Bc, max_partitions=Lc // Bc, n_inputs=nch, ring_len=eng.ring_length, fractional_delay=True,
                        nthreads=host_cores())
    obank = [[orc.filter(irs[c][k], Bc) for k in range(nbank)] for c in range(nch)]
    rng = np.random.default_rng(1004)
    nsw = 8
    xs = rng.uniform(-0.25, 0.25, (nsw * Tmax * Bc, nch)).astype(np.float32)
    s24 = np.zeros(xs.size * 3, dtype=np.uint8)
    orc.transfer(xs.view(np.uint8).reshape(-1), cpulibs.FMT_FLOAT, 0, 0, nch, s24, cpulibs.FMT_24, 0, 0, nch, nch, xs.shape[0])
    pos, gout, oout = 0, [], []
    for m in range(nsw):
        sel, dly, nb = c4_schedule(m, nch, nbank)
        eng.SelectFilters(range(nch), [bank[c][sel[c]] for c in range(nch)], delays=dly, crossfade=[m > 0] * nch)
        for c in range(nch):
            ocv.set_filter(c, obank[c][sel[c]], m > 0, dly[c])
        blk = s24[pos * nch * 3:(pos + nb * Bc) * nch * 3]
        gout.append(eng.Convolve(blk, bbx.FMT_24BIT, nch, bbx.FMT_24BIT, nch, nb * Bc).copy())
        oout.append(ocv.process(blk, cpulibs.FMT_24, nch, cpulibs.FMT_24, nch, nb * Bc).copy())
        pos += nb * Bc
    g, o = s24_to_float(np.concatenate(gout)), s24_to_float(np.concatenate(oout))
    par = compare(g, o)
    par["max_abs_lsb24"] = float(np.abs(g - o).max() * 2 ** 23)
    par["tolerance"] = "snr_db >= 110, max-abs <= 1 LSB(24) + 1e-5 x peak (int24 outputs, SURVEY.md 8.A)"
    par["ok"] = bool(par["snr_db"] >= 110.0 and np.abs(g - o).max() <= 2.0 ** -23 + 1e-5 * np.abs(o).max())
    par["against"] = "oracle convolver (CPU checker) on %d switching calls, %d frames x %d channels" % (nsw, pos, nch)
    # throughput: device-resident s24, one call per switch
    frames = Tmax * Bc
    xd = torch.randint(0, 255, (frames * nch * 3,), dtype=torch.uint8, device="cuda")
    yd = torch.empty(frames * nch * 3, dtype=torch.uint8, device="cuda")
    state = {"m": nsw, "blocks": 0}

    def step(i):
        sel, dly, nb = c4_schedule(state["m"], nch, nbank)
        state["m"] += 1
        eng.SelectFilters(range(nch), [bank[c][sel[c]] for c in range(nch)], delays=dly, crossfade=[True] * nch)
        eng.ConvolveDev(xd.data_ptr(), 2, nch, yd.data_ptr(), 2, nch, nb * Bc)
        if i >= 5:
            state["blocks"] += nb
    ms, launches = timed_steps(eng, step, steps)
    r = {"config": "C4: 32-channel dynamic IR switching, bank of 16 IRs/ch (4096 taps), select every 100 ms with crossfade + "
                   "per-channel fractional delay, 512-sample blocks, int24 in/out", "channels": nch,
         "value": nch * state["blocks"] * Bc / FS / (ms * 1e-3), "unit": "channel-s/s", "ms_per_step": ms / steps,
         "blocks_per_step": state["blocks"] / steps, "launches_per_step": launches / steps, "parity": par, "snr_db": par["snr_db"]}
    for c in range(nch):
        eng.SelectFilter(c, bank[c][0], delay=20.5)
    r["latency"] = block_latency(eng, bbx, 2, nch, 2, nch, Bc)
    eng.close()
    return r


def c5_inputs(frames, nin=64):
    return np.stack([make_noise(1000 + i, frames) for i in range(nin)], axis=1)


def c5_reference_window(x, outputs, n0, n1, nin=64, Lm=4096):
    """float64 direct convolution of the MIMO matrix rows `outputs` over all inputs, window [n0, n1)"""
    return {o: direct_window([(x[:, i], make_ir(2000 + 64 * o + i, Lm), 0, 1.0) for i in range(nin)], n0, n1) for o in outputs}


def leg_c5(bbx, torch, dev, steps, peaks):
    """BASELINE.json configs[4] (C5): 64-in x 64-out matrix of 4096-tap IRs, B = 512, 64-block steps.  The per-bin
    complex GEMM runs on the tensor cores (k_mimo_tc: tcgen05.mma kind::tf32, 3 MMAs per product for fp32 accuracy).
    achieved = TF32 MMA flops issued per launch / the kernel's average launch time (CUDA events around every launch);
    peak = the measured dense bf16 rate of MEASURED_PEAKS.json / 2 (TF32 runs at half the bf16 rate)."""
    nin = nout = 64
    Lm, Pm, Tm = 4096, 8, 64
    eng = bbx.Convolver(B, Pm, nin, n_outputs=nout, mode=bbx.MODE_MIMO, max_blocks=Tm, device=dev)
    for o in range(nout):
        for i in range(nin):
            eng.SelectFilter(o * nin + i, eng.CreateFilter(make_ir(2000 + 64 * o + i, Lm)))
    frames = Tm * B
    x = c5_inputs(frames)
    xd, yd = dev_bytes(torch, x), torch.empty(frames * nout * 4, dtype=torch.uint8, device="cuda")
    eng.ConvolveDev(xd.data_ptr(), 4, nin, yd.data_ptr(), 4, nout, frames)
    eng.Sync()
    y = yd.cpu().numpy().view(np.float32).reshape(frames, nout)
    n0, n1 = frames - 512, frames
    refs = c5_reference_window(x, (0, 63), n0, n1)
    par = worst([compare(y[n0:n1, o], refs[o]) for o in refs])
    for _ in range(5):
        eng.ConvolveDev(xd.data_ptr(), 4, nin, yd.data_ptr(), 4, nout, frames)
    eng.Sync()
    eng.profile_mac(True)
    l0 = eng.launch_count()
    eng.timer_start()
    for _ in range(steps):
        eng.ConvolveDev(xd.data_ptr(), 4, nin, yd.data_ptr(), 4, nout, frames)
    ms = eng.timer_stop()
    launches = eng.launch_count() - l0
    mac = eng.mac_time()
    eng.profile_mac(False)
    n_tc, status = eng.tensor_status()
    lms = mac["ms"] / max(1, mac["launches"])
    # per bin: M = 128 (64 outputs x re/im), K = 2 * nin * P = 1024, N = 64 block-steps, 3 TF32 MMAs per product
    flops = 3 * 2 * 128 * (2 * nin * Pm) * Tm * B
    peak = peaks["bf16_tflops"] / 2
    ach = flops / (lms * 1e-3) / 1e12 if lms > 0 else 0.0
    roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": 190.9e6,
            "traffic_source": "ncu --set full capture under profiles/ (dram__bytes_read + write of one launch), not this run",
            "kernel": "k_mimo_tc", "launch_ms": lms, "tf32_flops_per_launch": flops,
            "peak_source": peaks["bf16_src"] + " / 2 (TF32 = half the bf16 rate)",
            "tensor_launches": n_tc, "status": status, "mac_share_of_step": mac["ms"] / ms if ms > 0 else None,
            "useful_fp32_equivalent_TFLOPs": flops / 3 / (lms * 1e-3) / 1e12 if lms > 0 else 0.0}
    r = {"config": "C5: MIMO 64-in x 64-out 4096-tap convolution matrix (4096 paths), 512-sample blocks, per-bin complex GEMM on the "
                   "tensor cores, one GPU", "channels": nout, "paths": nin * nout,
         "value": nout * steps * frames / FS / (ms * 1e-3), "unit": "output-channel-s/s", "ms_per_step": ms / steps, "blocks_per_step": Tm,
         "launches_per_step": launches / steps, "parity": par, "snr_db": par["snr_db"],
         "latency": block_latency(eng, bbx, 4, nin, 4, nout, B, n=4000)}
    eng.close()
    return r, roof


# ---- N > 1: the multi-GPU paths of SURVEY.md 8(e) ----------------------------------------------------------------------
def leg_c5_sharded(bbx, torch, dist, rank, world, dev, mode, steps):
    """C5 over `world` ranks.  mode "peer" / "nccl": INPUT-sharded (rank g holds 64 / world inputs and every output of the
    matrix; partial output spectra are summed over the ranks by peer-memory NVLink stores + rank-order sums, or by one
    ncclReduceScatter per call; rank g converts 64 / world outputs).  mode "outputs": OUTPUT-sharded (rank g holds every
    input and 64 / world outputs: no collective).  Device time = max over ranks between barriers."""
    nin = nout = 64
    Lm, Pm, Tm = 4096, 8, 64
    frames = Tm * B
    x = c5_inputs(frames)
    o0, no = bbx.shard_range(nout, rank, world)
    comm = None
    if mode == "outputs":
        eng = bbx.Convolver(B, Pm, nin, n_outputs=no, mode=bbx.MODE_MIMO, max_blocks=Tm, device=dev)
        for oo in range(no):
            for i in range(nin):
                eng.SelectFilter(oo * nin + i, eng.CreateFilter(make_ir(2000 + 64 * (o0 + oo) + i, Lm)))
        xin, in_ch = x, nin
    else:
        i0, ni = bbx.shard_range(nin, rank, world)
        eng = bbx.Convolver(B, Pm, ni, n_outputs=nout, mode=bbx.MODE_MIMO, max_blocks=Tm, mimo_shard_world=world,
                            mimo_shard_rank=rank, device=dev)
        if mode == "peer":
            handles = [None] * world
            dist.all_gather_object(handles, eng.PeerExport())
            eng.PeerAttach(handles)
        else:
            uid = [bbx.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            comm = bbx.Comm(world, rank, uid[0], device=dev)
            eng.SetComm(comm)
        for o in range(nout):
            for i in range(ni):
                eng.SelectFilter(o * ni + i, eng.CreateFilter(make_ir(2000 + 64 * o + i0 + i, Lm)))
        xin, in_ch = np.ascontiguousarray(x[:, i0:i0 + ni]), ni
    xd, yd = dev_bytes(torch, xin), torch.empty(frames * no * 4, dtype=torch.uint8, device="cuda")

    def step(i=0):
        eng.ConvolveDev(xd.data_ptr(), 4, in_ch, yd.data_ptr(), 4, no, frames)

    # parity from the fresh state: rank 0's first and last output of the first call against the float64 direct convolution
    step()
    eng.Sync()
    par = None
    if rank == 0:
        y = yd.cpu().numpy().view(np.float32).reshape(frames, no)
        n0, n1 = frames - 512, frames
        refs = c5_reference_window(x, (o0, o0 + no - 1), n0, n1)
        par = worst([compare(y[n0:n1, o - o0], refs[o]) for o in refs])
    for _ in range(5):
        step()
    eng.Sync()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    eng.profile_mac(True)
    eng.timer_start()
    for _ in range(steps):
        step()
    ms = eng.timer_stop()
    dist.barrier()
    mac = eng.mac_time()
    xc = eng.exchange_time() if mode != "outputs" else None
    eng.profile_mac(False)
    n_tc, status = eng.tensor_status()
    t = torch.tensor([ms, mac["ms"] / max(1, mac["launches"]), (xc["ms"] / max(1, xc["exchanges"])) if xc else 0.0],
                     dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, mac_ms, xc_ms = [float(v) for v in t.tolist()]
    r = {"layout": {"peer": "input-sharded, peer-memory mixdown (NVLink stores into CUDA-IPC buffers + epoch flags, rank-order sums)",
                    "nccl": "input-sharded, one ncclReduceScatter (fp32 sum) per call",
                    "outputs": "output-sharded (every rank transforms all 64 inputs, owns 64 / N outputs), no collective"}[mode],
         "n_gpus": world, "value": nout * steps * frames / FS / (ms * 1e-3), "unit": "output-channel-s/s", "ms_per_step": ms / steps,
         "blocks_per_step": Tm, "steps": steps, "mac_launch_ms": mac_ms, "tensor_launches": n_tc, "status": status}
    if xc:
        sent = xc["bytes_sent"] / max(1, xc["exchanges"])
        r.update({"exchange_ms": xc_ms, "nvlink_bytes_sent_per_step_per_rank": sent,
                  "exchange_GBps_per_rank": sent / (xc_ms * 1e-3) / 1e9 if xc_ms > 0 else None,
                  "exchange": "k_gather_spectra_peer + k_peer_wait" if mode == "peer" else "k_gather_spectra + ncclReduceScatter",
                  "nvlink_reference": "770 GB/s per direction per GPU measured peer copy (B200_PROFILING.md)"})
    if par is not None:
        r["parity"], r["snr_db"] = par, par["snr_db"]
    dist.barrier()  # peer mode: nobody frees a receive buffer the others still have mapped
    eng.close()
    if comm is not None:
        comm.close()
    dist.barrier()
    return r


def leg_c3_strong(bbx, torch, dist, rank, world, dev, steps, nblk):
    """C3 strong scaling: the 128-channel renderer split over the ranks, 128 / N channels per GPU."""
    c0, nc = bbx.shard_range(NCH, rank, world)
    eng = bbx.Convolver(B, P, nc, max_blocks=nblk, device=dev)
    irs0 = make_ir(2000 + c0, L)
    for c in range(nc):
        eng.SelectFilter(c, eng.CreateFilter(irs0 if c == 0 else make_ir(2000 + c0 + c, L)))
    frames = nblk * B
    x = np.stack([make_noise(1000 + c0 + c, frames) for c in range(nc)], axis=1)
    xd, yd = dev_bytes(torch, x), torch.empty(frames * nc * 4, dtype=torch.uint8, device="cuda")
    hins = [bbx.PinnedBuffer(frames * nc * 4) for _ in range(2)]
    houts = [bbx.PinnedBuffer(frames * nc * 4) for _ in range(2)]
    for h in hins:
        h.array[:] = x.view(np.uint8).reshape(-1)

    def step(i=0):
        eng.ConvolveDev(xd.data_ptr(), 4, nc, yd.data_ptr(), 4, nc, frames)

    step()
    step()
    eng.Sync()
    par = None
    if rank == 0:
        y = yd.cpu().numpy().view(np.float32).reshape(frames, nc)
        n0, n1 = frames - 1024, frames
        xx = np.concatenate([x[:, 0], x[:, 0]])
        par = worst([compare(y[n0:n1, 0], direct_window([(xx, irs0, 0, 1.0)], frames + n0, frames + n1))])
    for _ in range(3):
        step()
    eng.Sync()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    eng.profile_mac(True)
    eng.timer_start()
    for _ in range(steps):
        step()
    ms = eng.timer_stop()
    dist.barrier()
    mac = eng.mac_time()
    eng.profile_mac(False)
    evict_cpu_caches()
    for i in range(4):
        eng.ConvolveHostPtrAsync(hins[i & 1].ptr, 4, nc, houts[i & 1].ptr, 4, nc, frames)
    eng.Sync()
    dist.barrier()
    eng.timer_start()
    for i in range(steps):
        eng.ConvolveHostPtrAsync(hins[i & 1].ptr, 4, nc, houts[i & 1].ptr, 4, nc, frames)
    ms_e2e = eng.timer_stop()
    dist.barrier()
    t = torch.tensor([ms, ms_e2e, mac["ms"] / max(1, mac["launches"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, mac_ms = [float(v) for v in t.tolist()]
    audio = NCH * frames / FS  # the whole 128-channel job per step
    r = {"scaling": "strong", "total_channels": NCH, "channels_per_gpu": nc, "n_gpus": world, "steps": steps, "blocks_per_step": nblk,
         "value": audio * steps / (ms * 1e-3), "unit": "channel-s/s", "ms_per_step": ms / steps, "mac_launch_ms": mac_ms,
         "e2e": {"value": audio * steps / (ms_e2e * 1e-3), "unit": "channel-s/s", "ms_per_step": ms_e2e / steps,
                 "h2d_bytes_per_step": frames * nc * 4, "d2h_bytes_per_step": frames * nc * 4},
         "rows_per_mac_cta": nc * P / 148.0}
    if par is not None:
        r["parity"], r["snr_db"] = par, par["snr_db"]
    eng.close()
    return r


def evict_cpu_caches(nbytes=1 << 30):
    """Push freshly written host buffers out of the CPU caches.  A pinned buffer whose lines are still dirty in the CPU's
    last-level cache is read by the copy engine through cache snoops: on the GPU boxes of this pool the H2D copy of such a
    buffer runs at 30..60 % of the rate it reaches once the lines have been written back (tools/pcie_numa.py: 0.53 ms
    against 0.306 ms per 16.8 MB, and the same buffer is fast a few seconds later).  The bench sends the same synthetic
    buffers every step, so their state is made definite -- resident in host DRAM -- before anything is timed."""
    a = np.zeros(nbytes, dtype=np.uint8)
    step = 4 << 20  # read-modify-write in pieces small enough that no library switches to cache-bypassing stores
    for i in range(0, nbytes, step):
        a[i:i + step] += 1
    s = int(a[::4096].sum())
    del a
    return s


def host_path_rate(torch, dist, nbytes, reps=30):
    """Full-duplex host<->device copy rate of this box with EVERY rank copying at once (plain pinned copies of one step's
    input and output on two streams, no engine involved): the bound of the e2e leg.  GB/s each way per rank, min over ranks."""
    hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    e0, e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def burst(n, timed):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        if timed:
            e0.record()
            s1.wait_event(e0)
            s2.wait_event(e0)
        for _ in range(n):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        if timed:
            e1.record(s1)
            e2.record(s2)
        torch.cuda.synchronize()
    hin.fill_(1)
    hout.fill_(2)
    evict_cpu_caches()
    burst(3, False)
    rate = 0.0
    for _ in range(3):  # best of three bursts: this is the denominator of a roofline
        burst(reps, True)
        ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
        rate = max(rate, nbytes * reps / (ms * 1e-3) / 1e9)
    if dist is not None:
        t = torch.tensor([rate], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        rate = float(t.item())
    return rate


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
    ap.add_argument("--no-mimo", action="store_true", help="skip the C5 MIMO legs (roofline_mimo, configs.C5, the sharded legs at N > 1)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C2 / C4 legs")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling C3 leg")
    ap.add_argument("--leg-steps", type=int, default=200, help="timed steps of the secondary legs")
    ap.add_argument("--channels", type=int, default=0, help="channels per GPU of the headline workload (tuning runs; default 128)")
    args = ap.parse_args()
    if args.channels:
        globals()["NCH"] = args.channels
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
    # and 4): the C3 path has no GPU-to-GPU traffic, and the host-to-device copies of the e2e leg then share fewer PCIe switch
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
    peaks = measured_peaks()

    # ---- set up the rank's shard: 128 channels, distinct IRs, noise resident in HBM ----
    eng = bbx.Convolver(B, P, NCH, max_blocks=nblk, device=local, mac_ctas_per_sm=args.occ,
                        mac_l2_keep_16ths=args.l2keep, mac_time_tile=args.tile)
    ir_first, ir_last = make_ir(2000 + NCH * rank, L), make_ir(2000 + NCH * rank + NCH - 1, L)
    for c in range(NCH):
        h = ir_first if c == 0 else (ir_last if c == NCH - 1 else make_ir(2000 + NCH * rank + c, L))
        eng.SelectFilter(c, eng.CreateFilter(h))
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

    # ---- parity of the timed configuration, outside the timed region: the engine's first two steps from its fresh state
    #      (the same 64-block batch twice), last 1024 frames of the first and the last channel against float64 direct ----
    step_dev()
    step_dev()
    eng.Sync()
    parity = None
    if rank == 0:
        xh = x_dev[:, [0, NCH - 1]].cpu().numpy()
        yh = y_dev[:, [0, NCH - 1]].cpu().numpy()
        n0, n1 = frames - 1024, frames
        res = []
        for k, h in enumerate((ir_first, ir_last)):
            xx = np.concatenate([xh[:, k], xh[:, k]])
            res.append(compare(yh[n0:n1, k], direct_window([(xx, h, 0, 1.0)], frames + n0, frames + n1)))
        parity = worst(res)
        parity["checked"] = ("channels 0 and %d of the benchmarked engine, frames %d..%d of its second %d-block step, against a float64 "
                             "direct convolution of the same input (numpy)" % (NCH - 1, n0, n1, nblk))
        parity["pin"] = "definition: BlockConvolver / Convolver / FFTW are absent from the reference tree, no BBC output exists"

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
    kernel_name = eng.mac_kernel_name()  # the MAC kernel the timed steps ran
    eng.profile_mac(False)
    ms = max_over_ranks(ms)
    audio_s = NCH * frames / FS  # channel-seconds per step per rank
    value = world * audio_s * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through bbx_process ----
    evict_cpu_caches()  # the synthetic input sits in host DRAM, not in the CPU's cache (see evict_cpu_caches)
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
    # the bound of the e2e leg: this box's full-duplex host<->device copy rate with every rank copying
    link = host_path_rate(torch, dist, in_bytes)
    bound_ms = in_bytes / (link * 1e9) * 1e3
    e2e_roofline = {"bound": "pcie (host<->device copies, full duplex, all %d rank(s) copying)" % world, "achieved": in_bytes / (ms_e2e / args.steps * 1e-3) / 1e9,
                    "peak": link, "unit": "GB/s each way per GPU", "frac": bound_ms / (ms_e2e / args.steps),
                    "peak_source": "measured in this run: plain pinned cudaMemcpyAsync of one step's input and output on two streams, "
                                   "min over ranks", "bound_ms_per_step": bound_ms,
                    "bound_value": world * audio_s / (bound_ms * 1e-3)}

    # ---- per-block latency, streaming T = 1 through the host API ----
    latency = None
    if not args.no_latency and rank == 0:
        latency = block_latency(eng, bbx, bbx.FMT_FLOAT, NCH, bbx.FMT_FLOAT, NCH, B, n=2000, warm=100)
        latency["mode"] += " (direct path: the PCM kernels read / write the pinned buffers over PCIe)"
        eng.set_direct_io(0)  # the same call through the staged copy-engine pipeline, for comparison
        st = block_latency(eng, bbx, bbx.FMT_FLOAT, NCH, bbx.FMT_FLOAT, NCH, B, n=1000, warm=100)
        eng.set_direct_io(1 << 20)
        latency["staged_p50_us"], latency["staged_p99_us"] = st["p50_us"], st["p99_us"]

    # ---- roofline of the dominant kernel of the timed region, CUDA events around every MAC launch ----
    units_per_launch = mac["channel_blocks"] / max(1, mac["launches"])
    mac_ms = mac["ms"] / max(1, mac["launches"])
    traffic_tb = traffic_stream = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "mac_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic_stream = tj["dram_bytes_per_channel_block"]
            traffic_tb = tj.get("tb_dram_bytes_per_channel_block")
            traffic_src = "ncu --set full captures recorded in profiles/mac_traffic.json (round %s), scaled to this launch; not measured in this run" % tj.get("round", "?")
        except Exception:
            pass
    batched = args.tile != 1
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    fp32_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12  # TFLOP/s: 148 SMs x 128 FMA lanes x 2 flop x max SM clock
    # context: the rate a pure packed-FMA kernel with the MAC's operand pattern reaches in this run (MEASURED_PEAKS.json has
    # no FP32 figure): burst = best isolated launch, sustained = 0.5 s back to back
    fp32_burst, fp32_sustained = bbx.probe_fp32_tflops(local, 0.5)
    if batched:
        tflops = FLOPS_PER_CHANNEL_BLOCK * units_per_launch / (mac_ms * 1e-3) / 1e12 if mac_ms > 0 else 0.0
        roofline = {"bound": "fp32", "achieved": tflops, "peak": fp32_nominal, "unit": "TFLOP/s", "frac": tflops / fp32_nominal,
                    "traffic": traffic_tb * units_per_launch if traffic_tb else None, "traffic_source": traffic_src,
                    "kernel": kernel_name, "launch_ms": mac_ms, "units_per_launch": units_per_launch,
                    "flops_per_launch": FLOPS_PER_CHANNEL_BLOCK * units_per_launch,
                    "peak_source": "nominal: 148 SM x 128 lanes x 2 flop x %.0f MHz (MEASURED_PEAKS.json has no FP32 figure)" % sm_max,
                    "peak_probe": fp32_burst, "frac_probe": tflops / fp32_burst if fp32_burst > 0 else None,
                    "peak_probe_sustained": fp32_sustained,
                    "peak_probe_source": "pure packed-FMA kernel (bbx_probe_fp32_tflops) in this run: best of 5 isolated launches / 0.5 s back to back",
                    "note": "time-batched MAC: every H[p] row is loaded once per 16 block-steps, 4 FMA per loaded byte -> bound by the "
                            "FP32 pipe, not HBM; the hbm roofline of the streaming kernel is in roofline_streaming",
                    "mac_share_of_step": mac["ms"] / ms if ms > 0 else None}
    else:
        achieved = BYTES_PER_CHANNEL_BLOCK * units_per_launch / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0
        dram = traffic_stream * units_per_launch / (mac_ms * 1e-3) / 1e9 if (traffic_stream and mac_ms > 0) else None
        roofline = {"bound": "hbm", "achieved": dram if dram else achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": (dram if dram else achieved) / peaks["hbm_gbs"], "frac_algorithmic": achieved / peaks["hbm_gbs"],
                    "traffic": traffic_stream * units_per_launch if traffic_stream else None, "traffic_source": traffic_src,
                    "kernel": "k_fdl_mac", "launch_ms": mac_ms, "units_per_launch": units_per_launch,
                    "algorithmic_bytes_per_launch": BYTES_PER_CHANNEL_BLOCK * units_per_launch, "peak_source": peaks["hbm_src"],
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
        alg = BYTES_PER_CHANNEL_BLOCK * upl / (lms * 1e-3) / 1e9 if lms > 0 else 0.0
        dram = traffic_stream * upl / (lms * 1e-3) / 1e9 if (traffic_stream and lms > 0) else None
        roofline_streaming = {"bound": "hbm", "achieved": dram if dram else alg, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": (dram if dram else alg) / peaks["hbm_gbs"],
                              "achieved_algorithmic": alg, "frac_algorithmic": alg / peaks["hbm_gbs"],
                              "traffic": traffic_stream * upl if traffic_stream else None, "traffic_source": traffic_src,
                              "kernel": "k_fdl_mac", "launch_ms": lms, "units_per_launch": upl,
                              "algorithmic_bytes_per_launch": BYTES_PER_CHANNEL_BLOCK * upl, "peak_source": peaks["hbm_src"],
                              "value_streaming": world * audio_s * ks / (ms_s * 1e-3), "steps": ks,
                              "mac_share_of_step": mac_s["ms"] / ms_s if ms_s > 0 else None,
                              "note": "achieved / frac = DRAM bytes per second (the ncu capture's dram__bytes of this kernel, 0.83 of the "
                                      "algorithmic bytes: L2 evict_last hints keep 3/16 of the lines resident); *_algorithmic = SURVEY.md "
                                      "8(d) bytes per second, which exceeds the copy peak because of that residency"}

    # the headline engine is no longer needed: free its 1 GB before the secondary legs
    eng.close()
    for h in hins + houts:
        h.close()
    del x_dev, y_dev

    # ---- BASELINE.json's other configs on one GPU (rank 0, N = 1) ----
    configs = None
    roofline_mimo = None
    floor = launch_sync_floor(torch) if (rank == 0 and world == 1 and not args.no_latency) else None
    if floor is not None:
        cf = bbx.probe_launch_sync(local)
        floor["c_p50_us"], floor["c_min_us"] = float(np.percentile(cf, 50)), float(cf.min())
        floor["c_what"] = "the same round trip from C inside libbbx (bbx_probe_launch_sync): the floor under latency.p50_us"
    if rank == 0 and world == 1:
        configs = {"C3": {"config": WORKLOAD, "channels": NCH, "value": value, "unit": "channel-s/s", "ms_per_step": ms / args.steps,
                          "blocks_per_step": nblk, "parity": parity, "snr_db": parity["snr_db"] if parity else None, "latency": latency}}
        if not args.no_configs:
            configs["C1"] = leg_c1(bbx, torch, local, args.leg_steps)
            configs["C2"] = leg_c2(bbx, torch, local, args.leg_steps)
            configs["C4"] = leg_c4(bbx, torch, local, args.leg_steps)
        if not args.no_mimo:
            configs["C5"], roofline_mimo = leg_c5(bbx, torch, local, min(args.leg_steps, 100), peaks)

    # ---- N > 1: strong scaling of C3 and the sharded MIMO paths ----
    strong = None
    mimo_sharded = None
    if world > 1:
        if not args.no_strong:
            strong = leg_c3_strong(bbx, torch, dist, rank, world, local, max(20, min(args.leg_steps, 200)), nblk)
        if not args.no_mimo and 64 % world == 0:
            mimo_sharded = {}
            for mode, key in (("peer", "input_sharded_peer"), ("nccl", "input_sharded_nccl"), ("outputs", "output_sharded")):
                if mode == "nccl" and not bbx.lib().bbx_comm_available():
                    mimo_sharded[key] = {"unavailable": "libnccl.so.2 not loadable"}
                    continue
                mimo_sharded[key] = leg_c5_sharded(bbx, torch, dist, rank, world, local, mode, 100)
            mimo_sharded["workload"] = "C5: MIMO 64 in x 64 out, 4096-tap matrix, B=512, 64-block steps, f32 in/out, over %d GPUs" % world

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
            "config": c3_config(world, nblk),
            "run": {"mac": kernel_name if batched else "streaming (k_fdl_mac)",
                    "parallelism": "channel-sharded x%d, no collective" % world,
                    "devices": "rank r on GPU %d*r of %d visible" % (stride, ndev),
                    "host_binding": ("rank 0 bound to %d CPUs local to its GPU" % len(numa)) if numa else "none"},
            "x_realtime_per_channel": value / (NCH * world),
            "parity_pin": PARITY_PIN, "parity": parity,
            "roofline": roofline, "roofline_streaming": roofline_streaming, "roofline_mimo": roofline_mimo, "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "channel-s/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": in_bytes,
                    "ms_per_step": ms_e2e / args.steps, "roofline": e2e_roofline},
            "gpu_launches": int(launches), "clocks": clocks, "latency": latency, "launch_sync_floor": floor, "configs": configs,
            "strong_scaling": strong, "mimo_sharded": mimo_sharded,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
