"""Calibration sweep (scratch): wide vs narrow warp geometry over call size and chunk length, one process."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sdft_b200 import SDFT


def main():
    torch.cuda.set_device(0)
    out_cap = 6 << 30
    buf = torch.empty(out_cap, dtype=torch.uint8, device="cuda")
    for m, fd, ch in [(512, "f64", 1), (1024, "f64", 1), (4096, "f64", 1), (2048, "f32", 1), (512, "f64", 16)]:
        fdb = 16 if fd == "f64" else 8
        for n in (1024, 4096, 16384, 65536):
            if ch * n * m * fdb > out_cap // 2:
                continue
            x = torch.rand(ch * n, device="cuda", dtype=torch.float32) * 2 - 1
            ring = max(1, min(8, out_cap // (ch * n * m * fdb)))
            for geo in ("wide", "narrow"):
                os.environ["SDFT_B200_GEO"] = geo
                for L in (0, 32, 64, 128, 256):
                    g = SDFT(m, "hann", 1, td="f32", fd=fd, channels=ch)
                    g._use_torch_stream()
                    if L:
                        g.set_chunk(L)
                    f = g._f("sdft_batch")
                    xp = ctypes.c_void_p(x.data_ptr())
                    ops = [ctypes.c_void_p(buf.data_ptr() + r * ch * n * m * fdb) for r in range(ring)]
                    reps = max(8, min(128, int(3e-3 / max(1e-6, ch * n * m * fdb / 5e12))))
                    for r in range(3):
                        f(g._h, n, xp, ops[r % ring])
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for r in range(reps):
                        f(g._h, n, xp, ops[r % ring])
                    e1.record()
                    torch.cuda.synchronize()
                    g._check()
                    us = e0.elapsed_time(e1) * 1e3 / reps
                    print(json.dumps({"m": m, "fd": fd, "ch": ch, "n": n, "geo": geo, "L": L, "us": round(us, 2),
                                      "GBps": round(ch * n * m * fdb / us / 1e3, 1)}), flush=True)
                    del g
    os.environ.pop("SDFT_B200_GEO", None)


if __name__ == "__main__":
    main()
