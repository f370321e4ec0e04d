#!/bin/bash
OUT=gpurun_out/v17; mkdir -p $OUT
timeout 900 python tools/geo_sweep_big.py > $OUT/geo_big.jsonl 2> $OUT/geo_big.err
python - <<'PY'
import json
from collections import defaultdict
t=defaultdict(dict)
for l in open("gpurun_out/v17/geo_big.jsonl"):
    d=json.loads(l); t[(d["m"],d["fd"],d["ch"],d["n"])][(d["geo"],d["L"])]=d["GBps"]
for k in sorted(t):
    print(k, " | ".join("%s: %s" % (geo, " ".join("%d:%.0f" % (L, t[k][(geo,L)]) for L in (0,128,256,512) if (geo,L) in t[k])) for geo in ("wide","narrow")))
PY
tail -3 $OUT/geo_big.err
