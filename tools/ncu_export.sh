#!/bin/bash
# ncu --set full capture of one kernel launch, exported on the box to CSV (raw + SASS source pages); the
# .ncu-rep itself is dropped because gpurun brings back at most 64 MiB.
# Usage: bash tools/ncu_export.sh <out prefix> <kernel regex> <skip> <command...>
PFX=$1; KRE=$2; SKIP=$3; shift 3
timeout 900 ncu --set full --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__thread_inst_executed.sum --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o $PFX "$@" > $PFX.log 2>&1
ncu -i $PFX.ncu-rep --page raw --csv > $PFX.raw.csv 2>/dev/null
ncu -i $PFX.ncu-rep --page source --csv --print-source sass > $PFX.sass.csv 2>/dev/null
rm -f $PFX.ncu-rep
