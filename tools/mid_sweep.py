"""Short and mid-size calls, serial vs streaming (sdft_b200_set_streaming): device time per call of back-to-back
analysis calls with device buffers (CUDA events), default geometry / chunk choice, rows into distinct tiles.
Usage (GPU box): python tools/mid_sweep.py > profiles/r02_mid_sweep.md"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from sdft_b200 import SDFT
    buf = torch.empty(40 << 30, dtype=torch.uint8, device="cuda")
    print("# Short and mid-size calls: serial vs streaming (B200, tools/mid_sweep.py)\n")
    print("Device time per call of back-to-back analysis calls on one plan (device buffers, CUDA events, rows into distinct")
    print("tiles), default geometry / chunk choice, hann, f32 time domain.  Calls above 2^28 bin-updates stay serial on a")
    print("streaming plan.\n")
    print("| m | FD | n per call | serial us | serial GB/s | streaming(8) us | streaming GB/s | HBM-time us @6454 |")
    print("|---|---|---|---|---|---|---|---|")
    for m, fd in ((512, "f64"), (1024, "f64"), (2048, "f32"), (4096, "f64"), (4096, "f32")):
        esz = 16 if fd == "f64" else 8
        for n in (1024, 4096, 16384, 65536, 262144):
            tile = n * m * esz
            calls = min(256, buf.numel() // tile)
            if calls < 2:
                continue
            x = torch.rand(calls * n, device="cuda") * 2 - 1
            res = []
            for depth in (1, 8):
                g = SDFT(m, "hann", 1, td="f32", fd=fd)
                g._use_torch_stream()
                g.set_streaming(depth)
                f = g._f("sdft_hops")

                def run():
                    f(g._h, calls, n, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(buf.data_ptr()), n * m)
                run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    run()
                e1.record()
                torch.cuda.synchronize()
                g._check()
                res.append(e0.elapsed_time(e1) * 1e-3 / 3 / calls)
            print("| %d | %s | %d | %.1f | %.0f | %.1f | %.0f | %.1f |" % (m, fd, n, res[0] * 1e6, tile / res[0] / 1e9, res[1] * 1e6,
                                                                       tile / res[1] / 1e9, tile / 6454e9 * 1e6), flush=True)


if __name__ == "__main__":
    main()
