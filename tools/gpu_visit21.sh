#!/bin/bash
OUT=gpurun_out/v21; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "randomized" > $OUT/pytest_random.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_random.log
tail -8 $OUT/pytest_random.log
# sanitizers on a small subset (slow under instrumentation)
cat > /tmp/san.py <<'PY'
import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from sdft_b200 import SDFT
rng = np.random.default_rng(5)
for fd, window, lat, geo in (("f64", "hann", 1.0, "wide"), ("f64", "blackman", 0.5, "narrow"), ("f32", "hamming", 0.5, "wide"), ("f32", "hann", 1.0, "narrow")):
    os.environ["SDFT_B200_GEO"] = geo
    g = SDFT(100, window, lat, td="f32", fd=fd)
    g.set_chunk(32)
    for n in (1, 7, 213, 1000, 650):
        x = rng.uniform(-1, 1, n).astype(np.float32)
        d = g.sdft(x); y = g.isdft(d); r = g.roundtrip(x)
    b = SDFT(64, window, lat, td="f32", fd=fd, channels=3)
    xb = rng.uniform(-1, 1, (3, 500)).astype(np.float32)
    b.sdft(xb); b.roundtrip(xb)
print("sanitizer workload done")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python /tmp/san.py > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python /tmp/san.py > $OUT/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?" >> $OUT/sanitizer_synccheck.log
for t in memcheck racecheck synccheck; do echo "== $t"; tail -5 $OUT/sanitizer_$t.log; done
