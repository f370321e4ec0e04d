#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python gpurun_scratch/single.py 2>&1 | head -8
timeout 300 python tools/stream_sweep.py --quick | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
