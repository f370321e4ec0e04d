#!/bin/bash
OUT=gpurun_out/v13; mkdir -p $OUT
export SDFT_B200_LIB=$PWD/sdft_b200/libsdft_b200_trace.so
for args in "--n 4096 --m 512" "--n 16384 --m 4096"; do
  echo "== $args"; timeout 200 python tools/trace_call.py $args 2>&1 | tail -14
done > $OUT/trace.txt 2>&1
cat $OUT/trace.txt
