#!/bin/bash
TAG=${1:-v6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bindings.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
R=$OUT/sweep.jsonl; : > $R
qb() { timeout 300 python tools/quick_bench.py "$@" >> $R 2>> $OUT/sweep.err; }
qb --hostrows pinned --n 65536 --m 4096 --reps 3
qb --hostrows pageable --n 65536 --m 4096 --reps 3
SDFT_B200_PAGEABLE=driver qb --hostrows pageable --n 65536 --m 4096 --reps 3
SDFT_B200_COPY_THREADS=4 qb --hostrows pageable --n 65536 --m 4096 --reps 3
SDFT_B200_COPY_THREADS=16 qb --hostrows pageable --n 65536 --m 4096 --reps 3
SDFT_B200_TILE_MB=32 qb --hostrows pageable --n 65536 --m 4096 --reps 3
cat $R | cut -c1-400
