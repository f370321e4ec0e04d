#!/bin/bash
# tests + one N=1 bench line
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
for k in ("value","ms_per_step","roofline","c3_time_sharded","c3_exact_f64","c4_channel_sharded"):
    print(k, json.dumps(d.get(k))[:1500])
print(json.dumps(d["other_configs"])[:1500])
PY
tail -5 $OUT/bench.err
