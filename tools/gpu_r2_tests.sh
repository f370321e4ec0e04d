#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 300 python tools/stream_sweep.py --quick > $OUT/stream_sweep_quick.md 2>&1; cat $OUT/stream_sweep_quick.md
timeout 600 python tools/mid_sweep.py > $OUT/mid_sweep.md 2>&1; cat $OUT/mid_sweep.md
