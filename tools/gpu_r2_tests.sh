#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row_pointer or single_sample" 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --no-cpu --no-sharded > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_quick.json"))
print(d["value"], d["roofline"]["frac"])
print(json.dumps(d["other_configs"]["row_pointer_variant"]))
print(json.dumps(d["other_configs"]["single_sample_calls"]))
print({k:v for k,v in d["e2e"].items() if k in ("bound","pcie_ceiling_GBps","frac_of_pcie")})
PY
tail -3 $OUT/bench_quick.err
