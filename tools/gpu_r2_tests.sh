#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -16 $OUT/pytest_gpu.log
SDFT_B200_LIB=$PWD/sdft_b200/libsdft_b200_trace.so python tools/trace_call.py --n 4096 --m 512 --calls 64 2>&1 | tee $OUT/trace_serial.txt
for G in wide narrow; do SDFT_B200_GEO=$G python tools/quick_bench.py --n 1048576 --m 4096 --fd f64 --roundtrip --reps 5; done
