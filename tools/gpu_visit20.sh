#!/bin/bash
OUT=gpurun_out/v20; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
R=$OUT/sweep.jsonl; : > $R
qb() { timeout 300 python tools/quick_bench.py "$@" >> $R 2>> $OUT/sweep.err; }
for md in fast strict; do
export SDFT_B200_F32=$md
qb --n 4194304 --m 2048 --fd f32 --window hann --latency 0.5 --reps 10
qb --n 1048576 --m 4096 --fd f32 --window hann --reps 10
qb --n 1048576 --m 4096 --fd f32 --window blackman --reps 10
qb --n 65536 --m 2048 --fd f32 --window hann --reps 20
qb --n 4194304 --m 2048 --fd f32 --window hann --latency 0.5 --reps 5 --roundtrip
done
python - <<'PY'
import json
for l in open("gpurun_out/v20/sweep.jsonl"):
    d=json.loads(l)
    print(d["mode"], d["m"], d["window"], d["n"], ("GB/s %.0f" % d["GBps"]) if "GBps" in d else "bu/s %.3g" % d["bin_updates_per_s"])
PY
