#!/bin/bash
OUT=gpurun_out/v36; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
R=$OUT/sweep.jsonl; : > $R
qb() { timeout 300 python tools/quick_bench.py "$@" >> $R 2>> $OUT/sweep.err; }
qb --n 1048576 --m 4096 --fd f64 --window hann --reps 8
qb --n 1048576 --m 4096 --fd f64 --window hann --reps 8 --roi 1024,512
qb --n 1048576 --m 4096 --fd f64 --window hann --reps 8 --roi 0,2048
python - <<'PY'
import json
for l in open("gpurun_out/v36/sweep.jsonl"):
    d=json.loads(l); print(d["mode"], d["m"], d.get("roi"), "ms %.3f" % d["ms"], "GB/s %.0f" % d["GBps"], "bu/s %.3g" % d["bin_updates_per_s"])
PY
