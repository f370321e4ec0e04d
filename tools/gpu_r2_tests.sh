#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 3 --no-cpu --no-sharded > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_quick.json"))
print(d["value"], d["roofline"]["frac"])
print(json.dumps(d["other_configs"]["single_sample_calls"]))
print(json.dumps(d["other_configs"]["config5_stream"])[:600])
PY
