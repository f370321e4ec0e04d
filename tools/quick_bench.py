"""Scratch timing helper (not the contract bench): device-resident analysis/synthesis throughput.

    python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --window hann [--synth]
    python tools/quick_bench.py --stream 4096 --calls 2048 --m 512          # many small calls on one plan
Prints one JSON line.  The library variant is chosen with SDFT_B200_LIB, the chunk with --chunk.
"""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdft_b200 import SDFT
from bench import ClockSampler


def event_time(fn, reps, warm=2):
    times = []
    for r in range(reps + warm):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.median(times[warm:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 18)
    ap.add_argument("--m", type=int, default=4096)
    ap.add_argument("--window", default="hann")
    ap.add_argument("--latency", type=float, default=1.0)
    ap.add_argument("--td", default="f32")
    ap.add_argument("--fd", default="f64")
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--synth", action="store_true")
    ap.add_argument("--roundtrip", action="store_true")
    ap.add_argument("--stream", type=int, default=0, help="samples per call; times --calls back-to-back calls")
    ap.add_argument("--calls", type=int, default=1024)
    ap.add_argument("--host", action="store_true", help="streaming with pinned host samples in / host samples out (roundtrip)")
    ap.add_argument("--roi", default="", help="first,count: rows hold only that region of interest")
    ap.add_argument("--hostrows", default="", help="sdft_n + isdft_n with HOST rows: 'pageable' (numpy) or 'pinned'")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    g = SDFT(a.m, a.window, a.latency, td=a.td, fd=a.fd, channels=a.channels)
    if a.chunk:
        g.set_chunk(a.chunk)
    g._use_torch_stream()
    rowlen = a.m
    if a.roi:
        first, count = (int(v) for v in a.roi.split(","))
        g.set_roi(first, count)
        rowlen = count
    tdt = torch.float32 if a.td == "f32" else torch.float64
    fdt = torch.complex64 if a.fd == "f32" else torch.complex128
    fdb = 16 if a.fd == "f64" else 8
    res = {"m": a.m, "window": a.window, "td": a.td, "fd": a.fd, "channels": a.channels, "chunk": a.chunk,
           "lib": os.path.basename(os.environ.get("SDFT_B200_LIB", "libsdft_b200.so"))}
    cs = ClockSampler(0)
    cs.start()

    if a.hostrows:
        n = a.n
        x = (np.random.default_rng(0).uniform(-1, 1, n)).astype(np.float32 if a.td == "f32" else np.float64)
        cdt = np.complex128 if a.fd == "f64" else np.complex64
        if a.hostrows == "pinned":
            rows_t = torch.empty((n, a.m), dtype=fdt).pin_memory()
            rows = rows_t.numpy()
        else:
            rows = np.empty((n, a.m), cdt)
            rows[:] = 0          # touch the pages once: first-touch faults are the caller's, not the library's
        y = np.empty(n, x.dtype)
        V = ctypes.c_void_p
        fa, fs = g._f("sdft_n"), g._f("isdft_n")
        ta, ts = [], []
        for r in range(a.reps + 1):
            t0 = time.perf_counter(); fa(g._h, n, x.ctypes.data_as(V), rows.ctypes.data_as(V)); t1 = time.perf_counter()
            fs(g._h, n, rows.ctypes.data_as(V), y.ctypes.data_as(V)); t2 = time.perf_counter()
            ta.append(t1 - t0); ts.append(t2 - t1)
        g._check()
        ta, ts = min(ta[1:]), min(ts[1:])
        res.update({"mode": "hostrows-" + a.hostrows, "n": n, "pageable_path": os.environ.get("SDFT_B200_PAGEABLE", "staged"),
                    "sdft_n_GBps": n * a.m * fdb / ta / 1e9, "isdft_n_GBps": n * a.m * fdb / ts / 1e9,
                    "sdft_n_bin_updates_per_s": n * a.m / ta})
    elif a.stream:
        n, calls, ch = a.stream, a.calls, a.channels
        if a.host:
            x = torch.rand(calls, ch * n, dtype=tdt).pin_memory() * 2 - 1
            y = torch.empty(calls, ch * n, dtype=tdt).pin_memory()
            f = g._f("roundtrip_n")

            def run():
                for c in range(calls):
                    f(g._h, n, ctypes.c_void_p(x[c].data_ptr()), ctypes.c_void_p(y[c].data_ptr()))
        else:
            x = torch.rand(calls, ch * n, device="cuda", dtype=tdt) * 2 - 1
            tile_bytes = ch * n * a.m * fdb
            ring = max(2, min(calls, (2 << 30) // tile_bytes))          # >= 2 GiB of output tiles: not L2-resident
            out = torch.empty((ring, ch * n, a.m), dtype=fdt, device="cuda")
            f = g._f("sdft_batch")
            xp = [ctypes.c_void_p(x[c].data_ptr()) for c in range(calls)]
            op = [ctypes.c_void_p(out[c % ring].data_ptr()) for c in range(calls)]

            def run():
                for c in range(calls):
                    f(g._h, n, xp[c], op[c])
        run()
        torch.cuda.synchronize()
        g._check()
        t0 = time.perf_counter()
        ms = event_time(run, a.reps, warm=1)
        wall = (time.perf_counter() - t0) / (a.reps + 1)
        res.update({"mode": "stream-host-roundtrip" if a.host else "stream-device", "n_per_call": n, "calls": calls,
                    "us_per_call": ms * 1e3 / calls, "wall_us_per_call": wall * 1e6 / calls,
                    "bin_updates_per_s": ch * n * a.m * calls / (ms * 1e-3),
                    "GBps": ch * n * a.m * fdb * calls / (ms * 1e-3) / 1e9})
        g._check()
    else:
        n, ch = a.n, a.channels
        x = torch.rand(ch * n, device="cuda", dtype=tdt) * 2 - 1
        xp = ctypes.c_void_p(x.data_ptr())
        if a.roundtrip:
            y = torch.empty_like(x)
            yp = ctypes.c_void_p(y.data_ptr())
            ms = event_time(lambda: g._f("roundtrip_n")(g._h, n, xp, yp), a.reps)
            res.update({"mode": "roundtrip-device", "n": n, "ms": ms, "bin_updates_per_s": ch * n * a.m / (ms * 1e-3),
                        "samples_per_s": ch * n / (ms * 1e-3)})
        else:
            out = torch.empty((ch * n, rowlen), dtype=fdt, device="cuda")
            op = ctypes.c_void_p(out.data_ptr())
            ms = event_time(lambda: g._f("sdft_batch")(g._h, n, xp, op), a.reps)
            res.update({"mode": "analysis", "n": n, "ms": ms, "bin_updates_per_s": ch * n * a.m / (ms * 1e-3),
                        "GBps": ch * n * rowlen * fdb / (ms * 1e-3) / 1e9, "roi": a.roi})
            if a.synth:
                y = torch.empty(ch * n, device="cuda", dtype=tdt)
                yp = ctypes.c_void_p(y.data_ptr())
                ms2 = event_time(lambda: g._f("isdft_batch")(g._h, n, op, yp), a.reps)
                res.update({"synth_ms": ms2, "synth_samples_per_s": ch * n / (ms2 * 1e-3),
                            "synth_GBps": ch * n * a.m * fdb / (ms2 * 1e-3) / 1e9})
        g._check()
    res["clocks"] = cs.stop()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
