"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): both warp geometries, all emit modes
(rows, state only, fused synthesis), batch plans, multi-call state, and the round-2 paths: streaming mode (overlapping
calls, counter hand-over), rows after a fused round trip (mirror cells), row-pointer variants with scattered device
rows, a float call split into wide body + narrow tail, double plans on the roots-of-unity phase source, a float
plan on a coarse phase-table stride.  Run on a GPU box:
    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sdft_b200 import SDFT

rng = np.random.default_rng(5)
for fd, window, lat, geo in (("f64", "hann", 1.0, "wide"), ("f64", "blackman", 0.5, "narrow"),
                             ("f32", "hamming", 0.5, "wide"), ("f32", "hann", 1.0, "narrow")):
    os.environ["SDFT_B200_GEO"] = geo
    g = SDFT(100, window, lat, td="f32", fd=fd)
    g.set_chunk(32)
    for n in (1, 7, 213, 1000, 650):
        x = rng.uniform(-1, 1, n).astype(np.float32)
        d = g.sdft(x)
        y = g.isdft(d)
        r = g.roundtrip(x)
        g.sdft(x[: max(1, n // 2)])          # rows right after a fused round trip
        g.advance(x)
    b = SDFT(64, window, lat, td="f32", fd=fd, channels=3)
    xb = rng.uniform(-1, 1, (3, 500)).astype(np.float32)
    b.sdft(xb)
    b.roundtrip(xb)
    # streaming: overlapping calls of every size on device buffers, a serial call in between
    s = SDFT(128, window, lat, td="f32", fd=fd)
    s.set_streaming(4)
    xt = torch.from_numpy(rng.uniform(-1, 1, 6000).astype(np.float32)).cuda()
    pos = 0
    for n in (64, 64, 300, 1, 1000, 64, 64, 2000, 17, 64, 64, 64):
        s.sdft(xt[pos:pos + n])
        pos += n
        if n == 17:
            s.roundtrip(xt[:500])
    s.advance(xt[:700])
    s.synchronize()
os.environ.pop("SDFT_B200_GEO", None)

# row pointers scattered through device memory
m, n = 96, 700
g = SDFT(m, "hann", 1, td="f32", fd="f64")
dev = torch.zeros((n, m), dtype=torch.complex128, device="cuda")
perm = rng.permutation(n)
ptrs = (ctypes.c_void_p * n)(*[dev.data_ptr() + int(p) * m * 16 for p in perm])
x = rng.uniform(-1, 1, n).astype(np.float32)
g._lib.sdft_b200_f32f64_sdft_nd(g._h, n, x.ctypes.data_as(ctypes.c_void_p), ptrs)
y = np.zeros(n, np.float32)
g._lib.sdft_b200_f32f64_isdft_nd(g._h, n, ptrs, y.ctypes.data_as(ctypes.c_void_p))
g._check()

# a long float call: wide body + narrow tail of bins in one launch (m = 1024: 4 full wide groups + 32 bins)
g = SDFT(1024, "hann", 0.5, td="f32", fd="f32")
xt = torch.from_numpy(rng.uniform(-1, 1, 330000).astype(np.float32)).cuda()
rows = g.sdft(xt)
assert g._lib.sdft_b200_split_count(g._h) == 1
g.advance(xt)
del rows

# coarse float phase table (stride 1024) and a big double plan (roots of unity), calls off the table grid
os.environ["SDFT_B200_F0_BUDGET_MB"] = "0"
g = SDFT(1000, "blackman", 0.5, td="f32", fd="f32")
os.environ.pop("SDFT_B200_F0_BUDGET_MB")
for n in (77, 1500, 333):
    g.sdft(rng.uniform(-1, 1, n).astype(np.float32))
g = SDFT(16384, "hamming", 1, td="f64", fd="f64")
for n in (5, 300):
    g.sdft(rng.uniform(-1, 1, n))
torch.cuda.synchronize()
print("sanitizer workload done")
