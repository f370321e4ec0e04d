#!/bin/bash
# One GPU visit: parity tests, smoke, both bench arms, ncu launch list of the bench command, full ncu captures of
# the dominant kernels, the streaming / mid-size sweeps.  Usage (under gpurun): bash tools/gpu_round.sh <tag>
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
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-sharded > $OUT/bench_under_ncu.log 2>&1
bash tools/ncu_export.sh $OUT/prof_f64 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1
bash tools/ncu_export.sh $OUT/prof_f32 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f32 --reps 1
bash tools/ncu_export.sh $OUT/prof_fused_f64 scan_emit 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1 --roundtrip
bash tools/ncu_export.sh $OUT/prof_synth_f64 synth_kernel 2 python tools/quick_bench.py --n 262144 --m 4096 --fd f64 --reps 1 --synth
bash tools/ncu_export.sh $OUT/prof_stream_call scan_emit 40 python tools/quick_bench.py --stream 4096 --calls 64 --m 512 --fd f64
M=sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
for D in 1 8; do
  timeout 300 ncu --replay-mode range --metrics $M --clock-control none --csv --log-file $OUT/stream_range_d$D.csv python tools/stream_range.py --depth $D > $OUT/stream_range_d$D.log 2>&1
done
timeout 600 python tools/mid_sweep.py > $OUT/mid_sweep.md 2>&1
timeout 600 python tools/stream_sweep.py > $OUT/stream_sweep.md 2>&1
tail -12 $OUT/pytest_gpu.log; cat $OUT/smoke.log | tail -3; cut -c1-3000 $OUT/bench.json; tail -3 $OUT/bench.err
