#!/bin/bash
TAG=${1:-v3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bindings.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 1500 python -m pytest tests/test_gpu_configs.py -m gpu -q -s --durations=10 > $OUT/pytest_configs.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_configs.log
tail -25 $OUT/pytest_configs.log
timeout 900 python tools/chunk_sweep.py > $OUT/chunk_sweep.jsonl 2> $OUT/chunk_sweep.err
R=$OUT/sweep.jsonl; : > $R
qb() { timeout 300 python tools/quick_bench.py "$@" >> $R 2>> $OUT/sweep.err; }
qb --n 1048576 --m 4096 --fd f64 --window hann --reps 12
qb --n 1048576 --m 4096 --fd f64 --window blackman --reps 12
qb --n 1048576 --m 4096 --fd f32 --window hann --reps 12
qb --n 1048576 --m 4096 --fd f64 --window hann --reps 12 --chunk 1024
qb --stream 4096 --calls 2048 --m 512 --fd f64 --reps 3
qb --stream 4096 --calls 2048 --m 512 --fd f64 --reps 3 --channels 16
qb --stream 4096 --calls 512 --m 512 --fd f64 --reps 3 --host
qb --n 1048576 --m 4096 --fd f64 --window hann --reps 5 --roundtrip
cat $R | cut -c1-300
