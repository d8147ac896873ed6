#!/bin/bash
# Multi-GPU visit: bench.py under torchrun at N = the box's GPU count (cfg3 and cfg4), and the single-process device
# group (tools/time_group.py).   Usage (through gpurun --gpus N): bash tools/scale_round.sh <N> <tag>
set -u
N=${1:-8}; TAG=${2:-r1}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29611 bench.py --gpus $N --no-supplied > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench n$N rc=$?"; cat $OUT/bench_n${N}_$TAG.json | cut -c1-400
$TR --master-port 29612 bench.py --gpus $N --workload cfg4 --no-supplied --no-reduced > $OUT/bench_cfg4_n${N}_$TAG.json 2> $OUT/bench_cfg4_n${N}_$TAG.err; echo "bench cfg4 n$N rc=$?"; cat $OUT/bench_cfg4_n${N}_$TAG.json | cut -c1-400
python tools/time_group.py > $OUT/group_n${N}_$TAG.jsonl 2> $OUT/group_n${N}_$TAG.err; echo "group rc=$?"; cat $OUT/group_n${N}_$TAG.jsonl
