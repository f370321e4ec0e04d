#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -a "config 3 at 2^26\|time shards, f32" $OUT/pytest_gpu.log
tail -16 $OUT/pytest_gpu.log
timeout 300 python tools/stream_sweep.py --quick
