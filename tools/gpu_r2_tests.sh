#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 300 python tools/stream_sweep.py --quick > $OUT/stream_sweep_quick.md 2>&1; cat $OUT/stream_sweep_quick.md
M=sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
for D in 1 8; do
  timeout 300 ncu --replay-mode range --metrics $M --clock-control none --csv --log-file $OUT/stream_range_d$D.csv python tools/stream_range.py --depth $D > $OUT/stream_range_d$D.log 2>&1
  tail -3 $OUT/stream_range_d$D.log; tail -8 $OUT/stream_range_d$D.csv
done
timeout 600 python bench.py --steps 3 --no-cpu --no-sharded > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_quick.json"))
print(d["value"], d["roofline"]["frac"], d["roofline"]["frac_store_peak"])
print(json.dumps(d["other_configs"], indent=1)[:3500])
PY
tail -5 $OUT/bench_quick.err
