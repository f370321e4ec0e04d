#!/bin/bash
OUT=gpurun_out/${1:-san}
mkdir -p $OUT
for T in memcheck racecheck synccheck; do
  echo "# compute-sanitizer --tool $T python tools/sanitizer_workload.py   (B200, round 2)" > $OUT/sanitizer_$T.txt
  timeout 1500 compute-sanitizer --tool $T python tools/sanitizer_workload.py 2>&1 | grep -v "^$" | tail -40 >> $OUT/sanitizer_$T.txt
  echo "$T exit ${PIPESTATUS[0]}" >> $OUT/sanitizer_$T.txt
  tail -4 $OUT/sanitizer_$T.txt
done
