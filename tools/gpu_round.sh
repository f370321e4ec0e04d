#!/bin/bash
# One GPU visit: parity tests, both bench arms, ncu launch list of the bench command, full ncu captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/bench_under_ncu.log 2>&1
bash tools/ncu_export.sh $OUT/prof_f64 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1
bash tools/ncu_export.sh $OUT/prof_f32 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f32 --reps 1
bash tools/ncu_export.sh $OUT/prof_fused_f64 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1 --roundtrip
bash tools/ncu_export.sh $OUT/prof_synth_f64 synth_kernel 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1 --synth
tail -12 $OUT/pytest_gpu.log; cat $OUT/smoke.log | tail -3; cat $OUT/bench.json; tail -3 $OUT/bench.err
