#!/bin/bash
OUT=gpurun_out/v16; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
SDFT_B200_GEO=narrow timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_narrow.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_narrow.log
tail -4 $OUT/pytest_narrow.log
timeout 900 python tools/geo_sweep.py > $OUT/geo_sweep.jsonl 2> $OUT/geo_sweep.err
python - <<'PY'
import json
from collections import defaultdict
t=defaultdict(dict)
for l in open("gpurun_out/v16/geo_sweep.jsonl"):
    d=json.loads(l); t[(d["m"],d["fd"],d["ch"],d["n"])][(d["geo"],d["L"])]=d["us"]
for k in sorted(t):
    print(k, " | ".join("%s: %s" % (geo, " ".join("%d:%.0f" % (L, t[k][(geo,L)]) for L in (0,32,64,128,256))) for geo in ("wide","narrow")))
PY
