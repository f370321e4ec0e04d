"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): both warp geometries, all emit modes
(rows, state only via isdft path, fused synthesis), batch plans, multi-call state.  Run on a GPU box:
    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

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
        g.advance(x)
    b = SDFT(64, window, lat, td="f32", fd=fd, channels=3)
    xb = rng.uniform(-1, 1, (3, 500)).astype(np.float32)
    b.sdft(xb)
    b.roundtrip(xb)
print("sanitizer workload done")
