#!/bin/bash
TAG=${1:-v10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
R=$OUT/sweep.jsonl; : > $R
qb() { timeout 300 python tools/quick_bench.py "$@" >> $R 2>> $OUT/sweep.err; }
for lib in "" _nopipe; do
  export SDFT_B200_LIB=$PWD/sdft_b200/libsdft_b200$lib.so
  qb --n 1048576 --m 4096 --fd f64 --window hann --reps 12
  qb --n 1048576 --m 4096 --fd f64 --window blackman --reps 12
  qb --stream 4096 --calls 2048 --m 512 --fd f64 --reps 3
  qb --stream 1024 --calls 2048 --m 1024 --fd f64 --reps 3
  qb --n 16384 --m 4096 --fd f64 --window hann --reps 20
  qb --n 65536 --m 1024 --fd f64 --window hann --reps 20
  qb --n 1048576 --m 4096 --fd f64 --window hann --reps 5 --roundtrip
done
unset SDFT_B200_LIB
python - <<'PY'
import json
for l in open("gpurun_out/v10/sweep.jsonl"):
    d=json.loads(l)
    print(d["lib"], d["mode"], d["m"], d["window"], d.get("n", d.get("n_per_call")), ("GB/s %.0f" % d["GBps"]) if "GBps" in d else "bu/s %.3g" % d["bin_updates_per_s"], ("us/call %.1f" % d["us_per_call"]) if "us_per_call" in d else "ms %.3f" % d["ms"])
PY
