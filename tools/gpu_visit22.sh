#!/bin/bash
OUT=gpurun_out/v22; mkdir -p $OUT
for c in 0 256 512 1024; do
  SDFT_B200_CHUNK=$c timeout 600 python bench.py --steps 20 --no-cpu --no-extras > $OUT/bench_chunk$c.json 2> $OUT/bench_chunk$c.err
done
SDFT_B200_WARPS=2 timeout 600 python bench.py --steps 20 --no-cpu --no-extras > $OUT/bench_w2.json 2> $OUT/bench_w2.err
SDFT_B200_WARPS=8 SDFT_B200_CHUNK=256 timeout 600 python bench.py --steps 20 --no-cpu --no-extras > $OUT/bench_w8c256.json 2> $OUT/bench_w8c256.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v22/bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "value %.4g"%d["value"], "achieved %.0f"%d["roofline"]["achieved"], "frac %.3f"%d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["power_w"])
    except Exception as e: print(f, e)
PY
