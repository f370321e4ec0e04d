"""Scratch timing helper (not the contract bench): device-resident analysis/synthesis throughput."""
import argparse
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdft_b200 import SDFT
from bench import ClockSampler


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 18)
    ap.add_argument("--m", type=int, default=4096)
    ap.add_argument("--window", default="hann")
    ap.add_argument("--latency", type=float, default=1.0)
    ap.add_argument("--td", default="f32")
    ap.add_argument("--fd", default="f64")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--synth", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    g = SDFT(a.m, a.window, a.latency, td=a.td, fd=a.fd)
    if a.chunk:
        g.set_chunk(a.chunk)
    x = (torch.rand(a.n, device="cuda", dtype=torch.float32 if a.td == "f32" else torch.float64) * 2 - 1)
    out = None
    times = []
    cs = ClockSampler(0)
    cs.start()
    for r in range(a.reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = g.sdft(x) if out is None else _again(g, x, out)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = float(np.median(times[2:]))
    clk = cs.stop()
    fdb = 16 if a.fd == "f64" else 8
    res = {"n": a.n, "m": a.m, "window": a.window, "fd": a.fd, "chunk": a.chunk, "ms": t,
           "bin_updates_per_s": a.n * a.m / (t * 1e-3), "GBps": a.n * a.m * fdb / (t * 1e-3) / 1e9, "clocks": clk}
    if a.synth:
        ts = []
        for r in range(a.reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = g.isdft(out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t2 = float(np.median(ts[2:]))
        res.update({"synth_ms": t2, "synth_samples_per_s": a.n / (t2 * 1e-3), "synth_GBps": a.n * a.m * fdb / (t2 * 1e-3) / 1e9})
    print(json.dumps(res))


def _again(g, x, out):
    g._use_torch_stream()
    g._f("sdft_batch")(g._h, x.shape[-1], ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()))
    g._check()
    return out


if __name__ == "__main__":
    main()
