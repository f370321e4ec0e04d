#!/usr/bin/env python
"""
bench.py -- headline benchmark of the sliding-DFT hot path (BASELINE.json: "analysis bin-updates/s +
synth samples/s at 1/2/4/8 B200, % HBM roofline").

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 2^20 samples of white noise, m = 4096 bins, float time domain /
double frequency domain, all four windows.  One STEP = the analysis (sdft_sdft_n) of the whole signal
once per window, i.e. 4 * 2^20 * 4096 bin-updates, with samples and the (n, m) output resident in HBM
(64 GiB written per window, so nothing is absorbed by the 126 MB L2 and no flush is needed).
At N > 1 every rank runs that same workload on its own channel (channel sharding, no data-path
collective): weak scaling, `value` = bin-updates of all ranks / max-over-ranks device time.

Extra keys on the JSON line:
  synthesis     sdft_isdft_n over the same matrix: samples/s and its own read roofline
  e2e           the same analysis metric through the C-ABI with HOST buffers (pinned), host->device and
                device->host copies inside the timed region (bounded sample, see e2e.sample)
  roofline      dominant kernel (analysis emit): algorithmic bytes / CUDA-event launch duration vs the
                measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own C path (oracle/_ref, else the oracle port) on the box's host cores
                (rank 0, N = 1 only)

--impl reference times that CPU path alone on the same metric/config (bounded sample per step).
Only the cpu_baseline / --impl reference legs touch oracle/; the measured product path never does.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "analysis_bin_updates_per_s"
UNIT = "bin-updates/s"
WINDOWS = ("boxcar", "hann", "hamming", "blackman")
M = 4096
N_SAMPLES = 1 << 20
SEED = 0x5DF70002


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 50 ms in the background; start() early (nvidia-smi needs ~100 ms to come
    up), mark() the timed region, stop() reports only the samples taken inside the marked region."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.file = None
        self.t0 = None
        self.t1 = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "power_w": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            self.file.flush()
            rows = [r.strip().split(",") for r in open(self.file.name) if r.strip()]
            os.unlink(self.file.name)
        except Exception:
            rows = []
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                ts = self._stamp(r[0])
                if self.t0 is not None and ts is not None and (ts < self.t0 or (self.t1 is not None and ts > self.t1 + 0.05)):
                    continue
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(self.NAMES, r[4:8]):
                    if "Active" in v and "Not" not in v:
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out.update({"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w": float(np.median(pw)),
                        "reasons": sorted(reasons), "samples": len(sm)})
        return out


# ------------------------------------------------------------------------------------------------
# CPU reference leg (the only place that touches oracle/)
# ------------------------------------------------------------------------------------------------
SYNTH = {}   # synthesis samples/s of the last cpu_reference_run: {"all": ..., "single": ...}


def cpu_reference_run(samples_per_window, threads, steps, warmup):
    """One channel per host thread through the reference's sdft_sdft_n (all four windows per step).
    Returns (bin-updates/s aggregate, seconds per step, kind, single-thread bin-updates/s)."""
    from oracle import cpu_reference
    rng = np.random.default_rng(SEED)
    plans, kind = [], "port"
    for t in range(threads):
        row = []
        for w in range(4):
            p, kind = cpu_reference("f32", "f64", M, w, 1.0, fast=True)
            row.append(p)
        plans.append(row)
    xs = [rng.uniform(-1, 1, samples_per_window).astype(np.float32) for _ in range(threads)]
    outs = [np.zeros((samples_per_window, M), np.complex128) for _ in range(threads)]

    def work(t):
        for w in range(4):
            p = plans[t][w]
            fn = p._fn("ref_sdft_n" if kind == "reference" else "sdft_n", None,
                       [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p])
            fn(p.h, samples_per_window, xs[t].ctypes.data_as(ctypes.c_void_p), outs[t].ctypes.data_as(ctypes.c_void_p))

    def step(nthreads):
        ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    for _ in range(warmup):
        step(threads)
    total = sum(step(threads) for _ in range(steps))
    per_step = total / steps
    agg = threads * 4 * samples_per_window * M / per_step
    single = 4 * samples_per_window * M / step(1)

    # synthesis (sdft_isdft_n) over the rows just produced: all threads, then one
    ys = [np.zeros(samples_per_window, np.float32) for _ in range(threads)]

    def synth(t):
        p = plans[t][1]
        fn = p._fn("ref_isdft_n" if kind == "reference" else "isdft_n", None,
                   [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p])
        fn(p.h, samples_per_window, outs[t].ctypes.data_as(ctypes.c_void_p), ys[t].ctypes.data_as(ctypes.c_void_p))

    def synth_step(nthreads):
        ts = [threading.Thread(target=synth, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    synth_step(threads)
    SYNTH["all"] = threads * samples_per_window / min(synth_step(threads), synth_step(threads))
    SYNTH["single"] = samples_per_window / min(synth_step(1), synth_step(1))
    return agg, per_step, kind, single


def cpu_build(kind):
    import oracle
    if kind == "reference":
        return "unmodified c/src/sdft/sdft.h, gcc -std=gnu99 -DSDFT_NO_COMPLEX_H %s (oracle/Makefile)" % oracle.fast_flags()
    return "oracle port, gcc -O2 -ffp-contract=off"


def cpu_hop_pattern(samples, threads):
    """The reference's own usage pattern (test/test.c:79-80): sdft_sdft_n followed by sdft_isdft_n on the
    same hop buffer, one channel per host thread, hann.  Returns analysis bin-updates/s (synthesis time included)."""
    from oracle import cpu_reference
    plans = [cpu_reference("f32", "f64", M, 1, 1.0, fast=True) for _ in range(threads)]
    kind = plans[0][1]
    rng = np.random.default_rng(SEED + 1)
    xs = [rng.uniform(-1, 1, samples).astype(np.float32) for _ in range(threads)]
    ys = [np.zeros(samples, np.float32) for _ in range(threads)]
    bufs = [np.zeros((samples, M), np.complex128) for _ in range(threads)]
    V = ctypes.c_void_p

    def work(t):
        p = plans[t][0]
        pre = "ref_" if kind == "reference" else ""
        a = p._fn(pre + "sdft_n", None, [V, ctypes.c_size_t, V, V])
        b = p._fn(pre + "isdft_n", None, [V, ctypes.c_size_t, V, V])
        a(p.h, samples, xs[t].ctypes.data_as(V), bufs[t].ctypes.data_as(V))
        b(p.h, samples, bufs[t].ctypes.data_as(V), ys[t].ctypes.data_as(V))

    def step():
        ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    step()
    dt = min(step(), step())
    return threads * samples * M / dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    spw = 4096
    agg, per_step, kind, single = cpu_reference_run(spw, threads, args.steps, args.warmup)
    sample = "%d host threads x 4 windows x %d samples per step (one channel per thread)" % (threads, spw)
    line = {
        "impl": "reference", "metric": METRIC, "value": agg, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: white noise, m=4096, f32 TD / f64 FD, four windows; reference C "
                               "sdft_sdft_n on host cores, bounded sample", "m": M, "sample": sample},
        "cpu_baseline": {"value": agg, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "build": cpu_build(kind),
                         "single_thread": single, "synthesis_samples_per_s": SYNTH.get("all"),
                         "synthesis_single_thread_samples_per_s": SYNTH.get("single")},
        "e2e": {"value": agg, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "hop_pattern": {"value": cpu_hop_pattern(2048, threads), "unit": UNIT,
                                "sample": "sdft_sdft_n + sdft_isdft_n per hop (test/test.c:79-80), hann, 2048-sample hops, "
                                          "%d host threads, one channel per thread" % threads}},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# short device-resident measurements of the other BASELINE configurations' shapes (N = 1 only)
# ------------------------------------------------------------------------------------------------
def other_configs(torch, SDFT, scratch, peak):
    """configs[2..4] are parity-test cases, not bench lines; these are their per-GPU shapes timed briefly
    (CUDA events, 3 warm-up + 5 timed calls, rows written into the 64 GiB scratch so nothing is L2-resident)."""
    raw = scratch.view(torch.uint8).view(-1)

    def timed(fn, reps=5):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    def ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    res = {}
    # config 3 shape: one time shard of the chirp, m = 2048, float FD, latency 0.5
    m, n = 2048, min(1 << 22, raw.numel() // (2048 * 8))
    g = SDFT(m, "hann", 0.5, td="f32", fd="f32")
    g._use_torch_stream()
    x = torch.rand(n, device=raw.device) * 2 - 1
    y = torch.empty_like(x)
    t_a = timed(lambda: g._f("sdft_n")(g._h, n, ptr(x), ptr(raw)))
    t_s = timed(lambda: g._f("isdft_n")(g._h, n, ptr(raw), ptr(y)))
    t_r = timed(lambda: g._f("roundtrip_n")(g._h, n, ptr(x), ptr(y)))
    g._check()
    res["config3_shard"] = {"workload": "n=%d, m=2048, f32 FD, hann, latency 0.5" % n,
                            "analysis_bin_updates_per_s": n * m / t_a, "analysis_GBps": n * m * 8 / t_a / 1e9,
                            "analysis_frac_of_hbm_peak": n * m * 8 / t_a / 1e9 / peak,
                            "synthesis_samples_per_s": n / t_s, "synthesis_GBps": n * m * 8 / t_s / 1e9,
                            "fused_roundtrip_bin_updates_per_s": n * m / t_r, "fused_roundtrip_samples_per_s": n / t_r}
    # config 4 shape: 64 channels per GPU, m = 1024, double FD
    m, ch = 1024, 64
    n = min(1 << 16, raw.numel() // (ch * m * 16))
    g = SDFT(m, "hann", 1, td="f32", fd="f64", channels=ch)
    g._use_torch_stream()
    x = torch.rand(ch * n, device=raw.device) * 2 - 1
    y = torch.empty_like(x)
    t_a = timed(lambda: g._f("sdft_batch")(g._h, n, ptr(x), ptr(raw)))
    t_r = timed(lambda: g._f("roundtrip_n")(g._h, n, ptr(x), ptr(y)))
    g._check()
    res["config4_per_gpu"] = {"workload": "64 channels x %d samples per call, m=1024, f64 FD, hann, one launch" % n,
                              "analysis_bin_updates_per_s": ch * n * m / t_a, "analysis_GBps": ch * n * m * 16 / t_a / 1e9,
                              "analysis_frac_of_hbm_peak": ch * n * m * 16 / t_a / 1e9 / peak,
                              "fused_roundtrip_bin_updates_per_s": ch * n * m / t_r}
    # config 5 shape: endless streaming in 4096-sample calls, m = 512, state carried across calls
    m, n, calls = 512, 4096, 1024
    x = torch.rand(calls * n, device=raw.device) * 2 - 1
    tile = n * m * 16
    xs = [ctypes.c_void_p(x.data_ptr() + c * n * 4) for c in range(calls)]
    os_ = [ctypes.c_void_p(raw.data_ptr() + c * tile) for c in range(calls)]

    def stream_leg(depth, from_c):
        g = SDFT(m, "hann", 1, td="f32", fd="f64")
        g._use_torch_stream()
        g.set_streaming(depth)
        f = g._f("sdft_n")
        hops = g._f("sdft_hops")

        def py_loop():
            for c in range(calls):
                f(g._h, n, xs[c], os_[c])

        def c_loop():
            hops(g._h, calls, n, xs[0], os_[0], n * m)
        t = timed(c_loop if from_c else py_loop, reps=3) / calls
        g._check()
        return t
    t_serial = stream_leg(1, False)
    t_serial_c = stream_leg(1, True)
    t_py = stream_leg(8, False)
    t_c = stream_leg(8, True)
    res["config5_stream"] = {"workload": "%d back-to-back calls of 4096 samples on one plan, m=512, f64 FD, hann; "
                                         "rows into %d distinct 32 MiB tiles; STREAMING mode "
                                         "(sdft_b200_set_streaming depth 8: consecutive calls overlap on the GPU), the hop "
                                         "loop issued from C (sdft_b200_*_sdft_hops = for h: sdft_sdft_n)" % (calls, calls),
                             "us_per_call": t_c * 1e6, "analysis_bin_updates_per_s": n * m / t_c,
                             "analysis_GBps": n * m * 16 / t_c / 1e9, "analysis_frac_of_hbm_peak": n * m * 16 / t_c / 1e9 / peak,
                             "us_per_call_streaming_python_loop": t_py * 1e6,
                             "us_per_call_serial_python_loop": t_serial * 1e6, "us_per_call_serial_c_loop": t_serial_c * 1e6,
                             "hbm_time_per_call_us": n * m * 16 / (peak * 1e9) * 1e6}
    # row-pointer variant against the contiguous call (sdft.h:622-628 vs :607-613), device rows, n = 2^16, m = 1024
    m, n = 1024, 1 << 16
    g = SDFT(m, "hann", 1, td="f32", fd="f64")
    g._use_torch_stream()
    x = torch.rand(n, device=raw.device) * 2 - 1
    rows_bytes = n * m * 16
    base = raw.data_ptr()
    contiguous = (ctypes.c_void_p * n)(*[base + i * m * 16 for i in range(n)])
    perm = np.random.default_rng(3).permutation(n)
    scattered = (ctypes.c_void_p * n)(*[base + int(perm[i]) * m * 16 for i in range(n)])
    t_n = timed(lambda: g._f("sdft_n")(g._h, n, ptr(x), ptr(raw)))
    t_nd = timed(lambda: g._f("sdft_nd")(g._h, n, ptr(x), contiguous))
    t_nds = timed(lambda: g._f("sdft_nd")(g._h, n, ptr(x), scattered))
    g._check()
    res["row_pointer_variant"] = {"workload": "sdft_sdft_nd vs sdft_sdft_n, n=65536, m=1024, f64 FD, device samples and rows: row "
                                              "pointers into one matrix (one run: the contiguous path) and in shuffled order "
                                              "(tile + one scatter kernel per tile)",
                                  "sdft_n_GBps": rows_bytes / t_n / 1e9, "sdft_nd_contiguous_GBps": rows_bytes / t_nd / 1e9,
                                  "sdft_nd_scattered_GBps": rows_bytes / t_nds / 1e9,
                                  "nd_contiguous_time_over_n": t_nd / t_n, "nd_scattered_time_over_n": t_nds / t_n}
    # the reference's per-sample entry points (sdft.h:562, :635) through the drop-in calls with host buffers
    m1 = 1000
    g = SDFT(m1, "hann", 1, td="f32", fd="f64")
    row = np.zeros(m1, np.complex128)
    rp = row.ctypes.data_as(ctypes.c_void_p)
    f_s, f_i = g._f("sdft"), g._f("isdft")
    for _ in range(50):
        f_s(g._h, ctypes.c_float(0.25), rp)
        f_i(g._h, rp)
    t0 = time.perf_counter()
    for k in range(500):
        f_s(g._h, ctypes.c_float(0.001 * k), rp)
    t_one = (time.perf_counter() - t0) / 500
    t0 = time.perf_counter()
    for k in range(500):
        f_i(g._h, rp)
    t_inv = (time.perf_counter() - t0) / 500
    g._check()
    res["single_sample_calls"] = {"workload": "sdft_sdft / sdft_isdft (ONE sample per call, host row), m=1000, f64 FD, hann; wall clock "
                                              "per call incl. the ctypes call; small calls travel through a pinned mailbox the "
                                              "kernels read and write in place", "sdft_us": t_one * 1e6, "isdft_us": t_inv * 1e6}
    return res


class Watchdog:
    """The sharded legs run collectives; a rank that dies inside one would leave the others waiting for NCCL's
    own timeout and the run without its line.  After `seconds` rank 0 prints the line it has (the legs
    marked as timed out) and every rank leaves."""

    def __init__(self, seconds, json_fd, partial_line):
        self.fd, self.line = json_fd, partial_line
        self.lock = threading.Lock()
        self.done = False
        self.timer = threading.Timer(seconds, self._bail)
        self.timer.daemon = True
        self.timer.start()

    def _bail(self):
        with self.lock:
            if self.done:
                return
            self.done = True
            if self.line is not None:
                os.write(self.fd, (json.dumps(self.line) + "\n").encode())
        os._exit(0)

    def cancel(self):
        with self.lock:
            self.done = True
        self.timer.cancel()


def bind_to_gpu_numa_node(gpu_index):
    """Pins this rank to the CPU cores next to its GPU (NVML's affinity mask), so that the pinned host
    buffers of the e2e legs are first-touched on the NUMA node the GPU's PCIe link hangs off."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, bits in enumerate(mask) for b in range(64) if (bits >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from sdft_b200 import SDFT

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sdft_b200 path has no CPU fallback")
    # stdout carries exactly ONE JSON line: libraries that print banners there (NCCL does) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, m = args.n, M
    free_b, _ = torch.cuda.mem_get_info()
    while n * m * 16 > 0.8 * free_b and n > 4096:
        n //= 2
    rng = np.random.default_rng([SEED, rank])
    x_host = rng.uniform(-1, 1, n).astype(np.float32)
    x = torch.from_numpy(x_host).to(dev)
    out = torch.empty((n, m), dtype=torch.complex128, device=dev)
    plans = [SDFT(m, w, 1, td="f32", fd="f64") for w in WINDOWS]
    stream_ptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for p in plans:
        p._lib.sdft_b200_set_stream(p._h, stream_ptr)
    xp, op = ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr())

    def analysis_step():
        for p in plans:
            p._f("sdft_n")(p._h, n, xp, op)

    y = torch.empty(n, dtype=torch.float32, device=dev)
    yp = ctypes.c_void_p(y.data_ptr())

    def synthesis_step():
        p = plans[-1]
        p._f("isdft_n")(p._h, n, op, yp)

    # ---- analysis: value + roofline --------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        analysis_step()
    for p in plans:
        p._check()
        p._lib.sdft_b200_set_profiling(p._h, 1)
    launches0 = sum(p.launches for p in plans)
    barrier()
    clocks.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        analysis_step()
    e1.record()
    torch.cuda.synchronize()
    clocks.mark_end()
    barrier()
    clk = clocks.stop()
    t_analysis = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = sum(p.launches for p in plans) - launches0
    kms, kcount = 0.0, 0
    for p in plans:
        c = ctypes.c_ulonglong(0)
        kms += p._lib.sdft_b200_kernel_ms(p._h, 0, ctypes.byref(c))
        kcount += c.value
        p._lib.sdft_b200_set_profiling(p._h, 0)
        p._check()
    value = world * args.steps * 4 * n * m / t_analysis
    peak, peak_src = measured_peaks()
    # ---- roofline denominators of THIS board in THIS run: pure store (the row kernel's own store instruction and
    # cache policy), pure read (the synthesis kernel's load), copy -- over 16 GiB of the row buffer (>> L2)
    lib = plans[0]._lib
    same_run = {}
    span = min(out.numel() * 16, 16 << 30)
    for kind, name in ((0, "store"), (1, "read"), (2, "copy")):
        sus = ctypes.c_double(0.0)
        best = lib.sdft_b200_measure_hbm(kind, ctypes.c_void_p(out.data_ptr()), span, 6, ctypes.byref(sus))
        same_run[name] = {"best_GBps": best, "mean_GBps": sus.value}
    dfma_peak = lib.sdft_b200_measure_dfma(5)
    launches += 3 * 8 + 7
    barrier()
    alg_bytes = n * m * 16 + n * 4
    dur = (kms / max(kcount, 1)) * 1e-3
    achieved = alg_bytes / dur / 1e9 if dur > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "emit_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_bin_update"] * n * m
        except Exception:
            traffic = None
    store_peak = same_run["store"]["best_GBps"] or None
    roofline = {"bound": "hbm", "kernel": "scan_emit_kernel<double> (chunk totals + look-back + row stores, one launch per call)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "static: profiles/emit_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu "
                                  "--set full capture of this kernel, per bin-update) x this launch's bin-updates",
                "peak_source": peak_src,
                "peak_store_same_run": store_peak, "frac_store_peak": (achieved / store_peak) if store_peak else None,
                "peak_read_same_run": same_run["read"]["best_GBps"], "peak_copy_same_run": same_run["copy"]["best_GBps"],
                "same_run_peaks": dict(same_run, how="sdft_b200_measure_hbm: pure streaming store (st.global"
                                       ".L1::no_allocate.L2::evict_first.v4.f64, the row kernel's instruction) / read / copy over "
                                       "%.0f GiB of device memory, best and mean of 6 launches, CUDA events" % (span / 2 ** 30)),
                "spec_GBps": 8000.0, "frac_spec": achieved / 8000.0,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dur * 1e3, "launches_timed": kcount,
                "whole_call_GBps": args.steps * 4 * alg_bytes / t_analysis / 1e9}

    # ---- synthesis -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        synthesis_step()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        synthesis_step()
    s1.record()
    torch.cuda.synchronize()
    barrier()
    t_synth = max_over_ranks(s0.elapsed_time(s1) * 1e-3)
    synth = {"metric": "synthesis_samples_per_s", "value": world * args.steps * n / t_synth, "unit": "samples/s",
             "ms_per_step": t_synth / args.steps * 1e3,
             "roofline": {"bound": "hbm", "achieved": args.steps * alg_bytes / t_synth / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": args.steps * alg_bytes / t_synth / 1e9 / peak,
                          "peak_read_same_run": same_run["read"]["best_GBps"],
                          "frac_read_peak": (args.steps * alg_bytes / t_synth / 1e9 / same_run["read"]["best_GBps"])
                          if same_run["read"]["best_GBps"] else None,
                          "spec_GBps": 8000.0, "frac_spec": args.steps * alg_bytes / t_synth / 1e9 / 8000.0}}
    del y
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            extras = other_configs(torch, SDFT, out, peak)
        except Exception as exc:      # the side measurements must never cost the headline line
            extras = {"error": "%s: %s" % (type(exc).__name__, exc)}
    del out
    torch.cuda.empty_cache()

    # ---- e2e: C-ABI call with host buffers (pinned), copies inside the timed region ---------------
    n_e = min(args.e2e_n, n)
    xe = torch.from_numpy(x_host[:n_e].copy()).pin_memory()
    oe = torch.empty((n_e, m), dtype=torch.complex128).pin_memory()
    # the ceiling of this leg on this box: plain pinned device->host copies of the same size, every rank at once
    src_e = torch.empty((n_e, m), dtype=torch.complex128, device=dev)
    oe.copy_(src_e, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        oe.copy_(src_e, non_blocking=True)
    torch.cuda.synchronize()
    t_pcie = max_over_ranks(time.perf_counter() - t0)
    barrier()
    pcie_gbps = 3 * n_e * m * 16 / t_pcie / 1e9
    del src_e
    pe = SDFT(m, "hann", 1, td="f32", fd="f64")
    xep, oep = ctypes.c_void_p(xe.data_ptr()), ctypes.c_void_p(oe.data_ptr())
    for _ in range(max(1, min(args.warmup, 3))):
        pe._f("sdft_n")(pe._h, n_e, xep, oep)
    pe._check()
    e2e_launch0 = pe.launches
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pe._f("sdft_n")(pe._h, n_e, xep, oep)      # returns when the rows are in host memory
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    pe._check()
    d2h_gbps = args.steps * n_e * m * 16 / t_e2e / 1e9
    e2e = {"value": world * args.steps * n_e * m / t_e2e, "unit": UNIT, "h2d_bytes_per_step": n_e * 4,
           "d2h_bytes_per_step": n_e * m * 16, "ms_per_step": t_e2e / args.steps * 1e3,
           "bound": "pcie",
           "bound_note": "the drop-in call delivers 16 B per bin-update to HOST memory; hop_pattern / roundtrip below "
                         "are what a caller gets who keeps the rows on the device",
           "d2h_GBps_per_gpu": d2h_gbps, "pcie_ceiling_GBps": pcie_gbps, "pcie_ceiling_GBps_per_gpu": pcie_gbps,
           "pcie_ceiling_GBps_all_gpus": pcie_gbps * world, "frac_of_pcie": d2h_gbps / pcie_gbps,
           "pcie_ceiling_how": "plain pinned cudaMemcpyAsync device->host of the same %.1f GiB buffer, %d rank(s) "
                               "concurrently, wall clock, max over ranks" % (n_e * m * 16 / 2 ** 30, world),
           "sample": "sdft_sdft_n(host pinned samples -> host pinned (n, m) rows), n=%d, m=%d, hann; "
                     "PCIe-bound: %.1f GB/s device->host" % (n_e, m, args.steps * n_e * m * 16 / t_e2e / 1e9)}
    launches += pe.launches - e2e_launch0
    del oe
    # the same call with PAGEABLE buffers (what malloc / NumPy callers of the reference hand over):
    # the library stages through its own pinned buffers and copies with host threads
    xg = x_host[:n_e].copy()
    og = np.zeros((n_e, m), np.complex128)
    xgp, ogp = xg.ctypes.data_as(ctypes.c_void_p), og.ctypes.data_as(ctypes.c_void_p)
    pe._f("sdft_n")(pe._h, n_e, xgp, ogp)
    pg_launch0 = pe.launches
    t0 = time.perf_counter()
    for _ in range(3):
        pe._f("sdft_n")(pe._h, n_e, xgp, ogp)
    t_pg = max_over_ranks(time.perf_counter() - t0)
    pe._check()
    launches += pe.launches - pg_launch0
    e2e["pageable"] = {"value": world * 3 * n_e * m / t_pg, "unit": UNIT,
                       "sample": "same call, pageable (NumPy) samples and rows: %.1f GB/s device->host through pinned "
                                 "staging + host copy threads" % (3 * n_e * m * 16 / t_pg / 1e9)}
    del og

    # the reference's own usage pattern (test/test.c:79-80): analysis then synthesis, hop by hop, with only
    # SAMPLES crossing PCIe (host in, host out) and the rows living in a device tile
    n_rt = min(1 << 20, n)
    xr = torch.from_numpy(x_host[:n_rt].copy()).pin_memory()
    yr = torch.empty(n_rt, dtype=torch.float32).pin_memory()
    pr = SDFT(m, "hann", 1, td="f32", fd="f64")
    xrp, yrp = ctypes.c_void_p(xr.data_ptr()), ctypes.c_void_p(yr.data_ptr())
    for _ in range(2):
        pr._f("roundtrip_n")(pr._h, n_rt, xrp, yrp)
    pr._check()
    rt_launch0 = pr.launches
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pr._f("roundtrip_n")(pr._h, n_rt, xrp, yrp)
    torch.cuda.synchronize()
    t_rt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    pr._check()
    launches += pr.launches - rt_launch0
    # the fused kernel alone (device samples in and out): FP64-issue bound; FP64 instructions per bin-update from
    # the committed ncu capture, the ceiling from a pure DFMA loop timed in this run
    xd_rt, yd_rt = xr.to(dev), torch.empty(n_rt, dtype=torch.float32, device=dev)
    fused_launch0 = pr.launches
    pr._use_torch_stream()
    for _ in range(2):
        pr._f("roundtrip_n")(pr._h, n_rt, ctypes.c_void_p(xd_rt.data_ptr()), ctypes.c_void_p(yd_rt.data_ptr()))
    f0_, f1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0_.record()
    for _ in range(args.steps):
        pr._f("roundtrip_n")(pr._h, n_rt, ctypes.c_void_p(xd_rt.data_ptr()), ctypes.c_void_p(yd_rt.data_ptr()))
    f1_.record()
    torch.cuda.synchronize()
    t_fused = f0_.elapsed_time(f1_) * 1e-3 / args.steps
    pr._check()
    fused_roofline = None
    fpath = os.path.join(ROOT, "profiles", "fused_fp64.json")
    if os.path.exists(fpath) and dfma_peak > 0:
        try:
            per = float(json.load(open(fpath))["fp64_thread_instructions_per_bin_update"])
            ach = per * n_rt * m / t_fused
            fused_roofline = {"bound": "fp64", "achieved": ach, "peak": dfma_peak, "unit": "FP64 thread-instructions/s",
                              "frac": ach / dfma_peak, "fp64_instructions_per_bin_update": per,
                              "device_resident_bin_updates_per_s": n_rt * m / t_fused,
                              "peak_source": "sdft_b200_measure_dfma: pure DFMA loop on every SM, this run",
                              "count_source": "static: profiles/fused_fp64.json (smsp__inst_executed_pipe_fp64.sum x 32 / "
                                              "bin-updates of one ncu capture of the fused kernel)"}
        except Exception:
            fused_roofline = None
    launches += pr.launches - fused_launch0
    del xd_rt, yd_rt
    e2e["roundtrip"] = {"roofline": fused_roofline, "value": world * args.steps * n_rt * m / t_rt, "unit": UNIT,
                        "samples_per_s": world * args.steps * n_rt / t_rt,
                        "h2d_bytes_per_step": n_rt * 4, "d2h_bytes_per_step": n_rt * 4,
                        "sample": "sdft_b200_f32f64_roundtrip_n(host samples -> host samples), n=%d, m=%d, hann: "
                                  "analysis + synthesis fused in one kernel, the rows never reach memory" % (n_rt, m)}

    # the same pattern through the DROP-IN calls: sdft_sdft_n(host samples -> hop buffer) and
    # sdft_isdft_n(hop buffer -> host samples), the hop buffer being device memory the caller allocated
    # instead of malloc; only samples cross PCIe
    hop = 1 << 16
    tile = torch.empty((hop, m), dtype=torch.complex128, device=dev)
    ph = SDFT(m, "hann", 1, td="f32", fd="f64")
    tp = ctypes.c_void_p(tile.data_ptr())

    def hop_run():
        for i in range(0, n_rt, hop):
            ph._f("sdft_n")(ph._h, hop, ctypes.c_void_p(xr.data_ptr() + 4 * i), tp)
            ph._f("isdft_n")(ph._h, hop, tp, ctypes.c_void_p(yr.data_ptr() + 4 * i))
    hop_run()
    ph._check()
    hp_launch0 = ph.launches
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hop_run()
    torch.cuda.synchronize()
    t_hp = max_over_ranks(time.perf_counter() - t0)
    barrier()
    ph._check()
    launches += ph.launches - hp_launch0
    e2e["hop_pattern"] = {"value": world * args.steps * n_rt * m / t_hp, "unit": UNIT,
                          "samples_per_s": world * args.steps * n_rt / t_hp,
                          "h2d_bytes_per_step": n_rt * 4, "d2h_bytes_per_step": n_rt * 4,
                          "sample": "sdft_sdft_n + sdft_isdft_n per %d-sample hop (test/test.c:79-80) on n=%d, m=%d, hann; host "
                                    "samples in and out, the caller's hop buffer is device memory" % (hop, n_rt, m)}
    del tile

    def make_line(cpu_leg):
        return {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_analysis / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: 2^20-sample white noise, m=4096, f32 TD / f64 FD, "
                                   "boxcar+hann+hamming+blackman per step; one channel per GPU",
                       "n_samples": n, "m": m, "windows": list(WINDOWS), "parallelism": "channel-sharded x%d" % world,
                       "l2": "no flush: each window writes %.0f GiB (>> 126 MB L2)" % (n * m * 16 / 2 ** 30)},
            "roofline": roofline, "synthesis": synth, "e2e": e2e, "cpu_baseline": cpu_leg, "clocks": clk,
            "other_configs": extras,
            "gpu_launches": int(launches),
        }

    # ---- the shardings BASELINE.json names: configs[2] by time, configs[3] by channel (every N) --------
    sharded = {}
    if not args.no_sharded:
        import bench_sharded
        partial = None
        if rank == 0:
            partial = dict(make_line(None), sharded_legs={"error": "timed out after %d s" % args.sharded_timeout})
        guard = Watchdog(args.sharded_timeout, json_fd, partial)
        sharded = bench_sharded.run_all(torch, dist if world > 1 else None, SDFT, dev, rank, world, reps=3)
        guard.cancel()
        launches += int(sum(v.get("gpu_launches_per_shard", 0) + v.get("gpu_launches_per_job", 0)
                            for v in sharded.values() if isinstance(v, dict)))

    # ---- CPU baseline beside it (rank 0, N = 1) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline gets every host core again
        threads = os.cpu_count() or 1
        spw = 4096
        agg, per_step, kind, single = cpu_reference_run(spw, threads, 2, 1)
        cpu = {"value": agg, "unit": UNIT, "cores": threads, "kind": kind, "single_thread": single, "build": cpu_build(kind),
               "synthesis_samples_per_s": SYNTH.get("all"), "synthesis_single_thread_samples_per_s": SYNTH.get("single"),
               "sample": "%d host threads x 4 windows x %d samples, one channel per thread, m=%d" % (threads, spw, m)}

    if rank == 0:
        line = make_line(cpu)
        line.update(sharded)
        line["gpu_launches"] = int(launches)
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_SAMPLES, help="samples per window (default 2^20)")
    ap.add_argument("--e2e-n", type=int, default=1 << 16, help="samples per e2e step (host buffers)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the short measurements of configs 3-5")
    ap.add_argument("--no-sharded", action="store_true", help="skip the time-/channel-sharded legs (configs 2 and 3)")
    ap.add_argument("--sharded-timeout", type=int, default=420, help="seconds after which the sharded legs are given up")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
