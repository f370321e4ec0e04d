#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 3 --no-cpu --no-sharded > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_quick.json"))
print(d["value"], d["roofline"]["frac"], d["roofline"]["frac_store_peak"])
print(json.dumps(d["other_configs"], indent=1)[:3500])
print(json.dumps(d["e2e"]["roundtrip"])[:1200])
PY
tail -5 $OUT/bench_quick.err
for NS in 0 1; do
  if [ $NS = 1 ]; then export SDFT_B200_NO_SPLIT=1; fi
  python tools/quick_bench.py --n 4194304 --m 2048 --fd f32 --latency 0.5 --reps 5
  python tools/quick_bench.py --n 2097152 --m 1024 --fd f32 --reps 5
done
