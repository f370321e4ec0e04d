#!/bin/bash
OUT=gpurun_out/v19; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
R=$OUT/sweep.jsonl; : > $R
qb() { timeout 300 python tools/quick_bench.py "$@" >> $R 2>> $OUT/sweep.err; }
qb --n 4194304 --m 2048 --fd f32 --window hann --latency 0.5 --reps 8 --synth
qb --n 4194304 --m 2048 --fd f32 --window hann --latency 1 --reps 8 --synth
qb --n 1048576 --m 4096 --fd f32 --window hann --reps 8 --synth
python - <<'PY'
import json
for l in open("gpurun_out/v19/sweep.jsonl"):
    d=json.loads(l)
    print(d["mode"], d["m"], d["fd"], d["n"], "GB/s %.0f" % d["GBps"], "synth GB/s %.0f" % d["synth_GBps"])
PY
