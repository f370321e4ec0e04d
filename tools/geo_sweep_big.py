"""Calibration sweep (scratch): wide vs narrow warp geometry for LARGE calls."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdft_b200 import SDFT

def main():
    torch.cuda.set_device(0)
    buf = torch.empty(68 << 30, dtype=torch.uint8, device="cuda")
    cases = [(4096, "f64", 1, 65536), (4096, "f64", 1, 262144), (4096, "f64", 1, 1048576), (4096, "f32", 1, 1048576),
             (1024, "f64", 64, 65536), (2048, "f32", 1, 4194304), (512, "f64", 1, 1048576)]
    for m, fd, ch, n in cases:
        fdb = 16 if fd == "f64" else 8
        x = torch.rand(ch * n, device="cuda", dtype=torch.float32) * 2 - 1
        for geo in ("wide", "narrow"):
            os.environ["SDFT_B200_GEO"] = geo
            for L in (0, 128, 256, 512):
                g = SDFT(m, "hann", 1, td="f32", fd=fd, channels=ch)
                g._use_torch_stream()
                if L:
                    g.set_chunk(L)
                f = g._f("sdft_batch")
                xp, op = ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(buf.data_ptr())
                for r in range(2):
                    f(g._h, n, xp, op)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 6
                e0.record()
                for r in range(reps):
                    f(g._h, n, xp, op)
                e1.record()
                torch.cuda.synchronize()
                g._check()
                us = e0.elapsed_time(e1) * 1e3 / reps
                print(json.dumps({"m": m, "fd": fd, "ch": ch, "n": n, "geo": geo, "L": L, "us": round(us, 1),
                                  "GBps": round(ch * n * m * fdb / us / 1e3, 1)}), flush=True)
                del g

if __name__ == "__main__":
    main()
