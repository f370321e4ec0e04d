#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the NCCL test and the bench at N ranks.
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi.log
tail -4 $OUT/pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_n$N.json"))
for k in ("value","ms_per_step","roofline","e2e","c3_time_sharded","c3_exact_f64","c4_channel_sharded","sharded_legs"):
    print(k, json.dumps(d.get(k))[:1800])
PY
tail -5 $OUT/bench_n$N.err
if [ "$3" != "noref" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
echo "ref exit $?"; cut -c1-300 $OUT/bench_ref_n$N.json
fi
