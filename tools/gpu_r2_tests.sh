#!/bin/bash
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for NS in 0 1; do
  export SDFT_B200_NO_SPLIT=$NS
  python tools/quick_bench.py --n 4194304 --m 2048 --fd f32 --latency 0.5 --reps 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no_split=$NS m=2048', round(d['GBps']))"
  python tools/quick_bench.py --n 2097152 --m 1024 --fd f32 --reps 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no_split=$NS m=1024', round(d['GBps']))"
done
unset SDFT_B200_NO_SPLIT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
