#!/bin/bash
OUT=gpurun_out/v15; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
R=$OUT/sweep.jsonl; : > $R
qb() { timeout 300 python tools/quick_bench.py "$@" >> $R 2>> $OUT/sweep.err; }
for pdl in 1 0; do
export SDFT_B200_PDL=$pdl
qb --stream 4096 --calls 2048 --m 512 --fd f64 --reps 3
qb --stream 1024 --calls 2048 --m 1024 --fd f64 --reps 3
qb --stream 4096 --calls 2048 --m 512 --fd f64 --reps 3 --channels 16
qb --stream 4096 --calls 512 --m 512 --fd f64 --reps 3 --host
qb --n 1048576 --m 4096 --fd f64 --window hann --reps 12
done
python - <<'PY'
import json
for l in open("gpurun_out/v15/sweep.jsonl"):
    d=json.loads(l)
    print(d["mode"], d["m"], d["fd"], d["channels"], d.get("n", d.get("n_per_call")), ("GB/s %.0f" % d["GBps"]) if "GBps" in d else "", ("us/call %.1f" % d["us_per_call"]) if "us_per_call" in d else "ms %.3f" % d["ms"])
PY
