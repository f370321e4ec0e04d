#!/bin/bash
# Profiling visit: bench (both arms), ncu launch list of the bench command, full ncu captures exported to CSV.
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/bench_under_ncu.log 2>&1
bash tools/ncu_export.sh $OUT/prof_f64 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1
bash tools/ncu_export.sh $OUT/prof_f32 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f32 --reps 1
bash tools/ncu_export.sh $OUT/prof_fused_f64 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1 --roundtrip
bash tools/ncu_export.sh $OUT/prof_synth_f64 synth_kernel 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1 --synth
cat $OUT/bench.json; tail -3 $OUT/bench.err; du -sh $OUT
