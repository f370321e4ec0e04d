"""Where does one analysis call spend its time?  Needs a -DSDFT_B200_TRACE build (SDFT_B200_LIB=...libsdft_b200_trace.so).
Prints, for the LAST of `--calls` back-to-back calls, the offsets (microseconds after the first CTA took its ticket)
of the per-CTA phase stamps: min / median / max over CTAs."""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdft_b200 import SDFT

NAMES = ["ticket", "deltas", "total", "agg_pub", "carry", "replay", "rows_done", "lb_seen"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--m", type=int, default=512)
    ap.add_argument("--fd", default="f64")
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--calls", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=0)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    g = SDFT(a.m, "hann", 1, td="f32", fd=a.fd, channels=a.channels)
    if a.chunk:
        g.set_chunk(a.chunk)
    g._use_torch_stream()
    fdb = 16 if a.fd == "f64" else 8
    x = torch.rand(a.calls, a.channels * a.n, device="cuda") * 2 - 1
    out = torch.empty((a.calls, a.channels * a.n * a.m * fdb), dtype=torch.uint8, device="cuda")
    f = g._f("sdft_batch")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        for c in range(a.calls):
            f(g._h, a.n, ctypes.c_void_p(x[c].data_ptr()), ctypes.c_void_p(out[c].data_ptr()))
        e1.record()
        torch.cuda.synchronize()
    g._check()
    print("device time per call: %.2f us" % (e0.elapsed_time(e1) * 1e3 / a.calls))
    buf = np.zeros((1 << 16, 8), np.uint64)
    items = g._lib.sdft_b200_debug_trace(g._h, buf.ctypes.data_as(ctypes.c_void_p), buf.shape[0])
    if not items:
        print("no trace: not a -DSDFT_B200_TRACE build")
        return
    t = buf[:items, :8].astype(np.int64)
    t0 = t[:, 0].min()
    print("CTAs: %d" % items)
    for k, name in enumerate(NAMES):
        col = (t[:, k] - t0) / 1e3
        col = col[t[:, k] > 0]
        if col.size:
            print("%-10s min %7.2f  median %7.2f  max %7.2f us" % (name, col.min(), np.median(col), col.max()))
    jb = np.arange(items) // max(1, items // max(1, (a.n + 63) // 64 // 4))   # rough block index (ticket order)
    lb = (t[:, 4] - t[:, 3]) / 1e3
    seen = np.where(t[:, 7] > 0, (t[:, 7] - t[:, 3]) / 1e3, 0)
    order = np.argsort(t[:, 3])
    print("look-back per CTA in ticket order (agg_pub->seen, seen->carry):")
    print(" ".join("%.1f/%.1f" % (seen[i], lb[i] - seen[i]) for i in range(0, items, max(1, items // 40))))
    span = (t[:, 6].max() - t0) / 1e3
    print("first ticket -> last rows done: %.2f us" % span)


if __name__ == "__main__":
    main()
