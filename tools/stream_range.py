"""One profiler range around K back-to-back 4096-sample calls at m = 512 (BASELINE config 5's shape), for
`ncu --replay-mode range`: with kernel replay ncu serialises the kernels, so the overlap of streaming calls -- the
whole point of the mode -- can only be seen over a range.  Usage (GPU box):
  ncu --replay-mode range --metrics <...> python tools/stream_range.py --depth 8"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--calls", type=int, default=256)
    ap.add_argument("--m", type=int, default=512)
    ap.add_argument("--hop", type=int, default=4096)
    a = ap.parse_args()
    from sdft_b200 import SDFT
    m, hop, calls = a.m, a.hop, a.calls
    x = torch.rand(calls * hop, device="cuda") * 2 - 1
    out = torch.empty((calls, hop, m), dtype=torch.complex128, device="cuda")
    g = SDFT(m, "hann", 1, td="f32", fd="f64")
    g._use_torch_stream()
    g.set_streaming(a.depth)
    f = g._f("sdft_hops")

    def run():
        f(g._h, calls, hop, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), hop * m)
    run()
    torch.cuda.synchronize()
    rt = torch.cuda.cudart()
    rt.cudaProfilerStart()
    run()
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    g._check()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    print("depth %d: %.2f us per call (outside the profiler range)" % (a.depth, e0.elapsed_time(e1) * 1e3 / calls))


if __name__ == "__main__":
    main()
