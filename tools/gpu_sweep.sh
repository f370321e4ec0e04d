#!/bin/bash
# A/B sweep of library variants and the non-headline configurations. Usage (under gpurun): bash tools/gpu_sweep.sh <tag>
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
R=$OUT/sweep.jsonl; : > $R
qb() { timeout 300 python tools/quick_bench.py "$@" >> $R 2>> $OUT/sweep.err; }
for lib in "" _mb5 _mb6 _h1; do
  export SDFT_B200_LIB=$PWD/sdft_b200/libsdft_b200$lib.so
  qb --n 1048576 --m 4096 --fd f64 --window hann --reps 12
  qb --n 1048576 --m 4096 --fd f64 --window blackman --reps 12
  qb --n 1048576 --m 4096 --fd f32 --window hann --reps 12
done
unset SDFT_B200_LIB
SDFT_B200_WARPS=4 qb --n 1048576 --m 4096 --fd f64 --window hann --reps 12
SDFT_B200_WARPS=4 qb --n 1048576 --m 4096 --fd f32 --window hann --reps 12
SDFT_B200_F32=strict qb --n 1048576 --m 4096 --fd f32 --window hann --reps 12
SDFT_B200_F64=modulated qb --n 1048576 --m 4096 --fd f64 --window hann --reps 12
qb --n 1048576 --m 4096 --fd f64 --window boxcar --reps 12 --synth
for c in 128 256 1024; do qb --n 1048576 --m 4096 --fd f64 --window hann --reps 8 --chunk $c; done
# config 3 shape (one shard): f32 FD, m=2048, latency 0.5
qb --n 4194304 --m 2048 --fd f32 --window hann --latency 0.5 --reps 8 --synth
# config 4 shape: 64 channels per GPU, m=1024, f64
qb --n 65536 --m 1024 --fd f64 --window hann --channels 64 --reps 8 --synth
# config 5 shape: streaming 4096-sample calls, m=512
qb --stream 4096 --calls 2048 --m 512 --fd f64 --reps 3
qb --stream 4096 --calls 2048 --m 512 --fd f64 --reps 3 --channels 16
qb --stream 4096 --calls 512 --m 512 --fd f64 --reps 3 --host
qb --stream 4096 --calls 1024 --m 1024 --fd f64 --reps 3
# round trip, device resident
qb --n 1048576 --m 4096 --fd f64 --window hann --reps 5 --roundtrip
cat $R | cut -c1-330
