"""Streaming mode (sdft_b200_set_streaming) tuning: microseconds per 4096-sample call at m = 512 (BASELINE config 5's
shape) over streaming depth, warp geometry, chunk length and CTA width.  Rows go to distinct 32 MiB tiles.
Usage (GPU box): python tools/stream_sweep.py [--m 512] [--hop 4096] [--calls 1024]"""
import argparse
import ctypes
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=512)
    ap.add_argument("--hop", type=int, default=4096)
    ap.add_argument("--calls", type=int, default=1024)
    ap.add_argument("--fd", default="f64")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    from sdft_b200 import SDFT
    m, hop, calls = a.m, a.hop, a.calls
    esz = 16 if a.fd == "f64" else 8
    x = torch.rand(calls * hop, device="cuda") * 2 - 1
    out = torch.empty(calls * hop * m * esz, dtype=torch.uint8, device="cuda")
    print("# Streaming mode tuning (B200, tools/stream_sweep.py)\n")
    print("Microseconds per %d-sample call at m = %d (%s frequency domain, hann), %d back-to-back calls issued by" % (hop, m, a.fd, calls))
    print("sdft_b200_*_sdft_hops into distinct tiles, over streaming depth (1 = serial), warp geometry, chunk length and")
    print("warps per CTA (auto = the library's choice).\n")
    print("| depth | geo | chunk | warps | us/call | GB/s |")
    print("|---|---|---|---|---|---|")
    depths = (1, 4, 8, 16) if not a.quick else (1, 8)
    geos = ("narrow", "wide")
    chunks = (0, 32, 64, 128, 256) if not a.quick else (0,)
    warps = (0, 2, 4, 8) if not a.quick else (0,)
    for depth, geo, chunk, w in itertools.product(depths, geos, chunks, warps):
        if chunk and w and chunk * w > 2048:
            continue
        os.environ["SDFT_B200_GEO"] = geo
        if w:
            os.environ["SDFT_B200_WARPS"] = str(w)
        else:
            os.environ.pop("SDFT_B200_WARPS", None)
        g = SDFT(m, "hann", 1, td="f32", fd=a.fd)
        g._use_torch_stream()
        g.set_streaming(depth)
        if chunk:
            g.set_chunk(chunk)
        f = g._f("sdft_hops")

        def run():
            f(g._h, calls, hop, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), hop * m)
        for _ in range(2):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        g._check()
        t = e0.elapsed_time(e1) * 1e-3 / 3 / calls
        print("| %d | %s | %s | %s | %.2f | %.0f |" % (depth, geo, chunk or "auto", w or "auto", t * 1e6, hop * m * esz / t / 1e9), flush=True)


if __name__ == "__main__":
    main()
