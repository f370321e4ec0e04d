import ctypes, json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from sdft_b200 import SDFT
torch.cuda.set_device(0)
out_cap = 6 << 30
buf = torch.empty(out_cap, dtype=torch.uint8, device="cuda")
os.environ["SDFT_B200_GEO"] = "narrow"
for m, fd, ch in [(512, "f64", 1), (1024, "f64", 1), (4096, "f64", 1), (2048, "f32", 1), (512, "f64", 16)]:
    fdb = 16 if fd == "f64" else 8
    for n in (1024, 4096, 16384, 65536):
        if ch * n * m * fdb > out_cap // 2:
            continue
        x = torch.rand(ch * n, device="cuda", dtype=torch.float32) * 2 - 1
        ring = max(1, min(8, out_cap // (ch * n * m * fdb)))
        row = {}
        for W in (2, 4, 8):
            os.environ["SDFT_B200_WARPS"] = str(W)
            for L in (32, 64, 128, 256):
                g = SDFT(m, "hann", 1, td="f32", fd=fd, channels=ch)
                g._use_torch_stream(); g.set_chunk(L)
                f = g._f("sdft_batch")
                xp = ctypes.c_void_p(x.data_ptr())
                ops = [ctypes.c_void_p(buf.data_ptr() + r * ch * n * m * fdb) for r in range(ring)]
                reps = max(8, min(128, int(3e-3 / max(1e-6, ch * n * m * fdb / 5e12))))
                for r in range(3): f(g._h, n, xp, ops[r % ring])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for r in range(reps): f(g._h, n, xp, ops[r % ring])
                e1.record(); torch.cuda.synchronize(); g._check()
                row[(W, L)] = e0.elapsed_time(e1) * 1e3 / reps
                del g
        best = min(row, key=row.get)
        print((m, fd, ch, n), " ".join("W%d:[%s]" % (W, " ".join("%.0f" % row[(W, L)] for L in (32, 64, 128, 256))) for W in (2, 4, 8)), "best", best, "%.0f" % row[best], flush=True)
