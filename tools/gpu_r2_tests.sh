#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 600 python tools/stream_sweep.py > $OUT/stream_sweep.md 2>&1; cat $OUT/stream_sweep.md
timeout 600 python bench.py --steps 3 --no-cpu --no-sharded > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_quick.json"))
print(d["value"], json.dumps(d["roofline"])[:1500])
print(json.dumps(d["synthesis"])[:800])
print(json.dumps(d["e2e"])[:2500])
print(json.dumps(d["other_configs"]["config5_stream"], indent=1)[:3000])
PY
tail -5 $OUT/bench_quick.err
