#!/bin/bash
# Multi-GPU visit for the peer-memory exchange of the cost sums: the multi-GPU checks, then bench.py (cfg3 and cfg4) with the
# sums meeting over peer memory (the default; also fused into the rollout kernel with "fuse") and in an ncclAllReduce.
#   Usage (through gpurun --gpus N): bash tools/peer_round.sh <N> <tag> [fuse] [group]
set -u
N=${1:-2}; TAG=${2:-r1}; shift 2
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29640 tools/multi_gpu_check.py > $OUT/multi_check_n${N}_$TAG.log 2>&1; echo "multi_gpu_check rc=$?"; grep "multi_gpu_check ok" $OUT/multi_check_n${N}_$TAG.log || tail -20 $OUT/multi_check_n${N}_$TAG.log
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "group or cpp" > $OUT/pytest_multi_n${N}_$TAG.log 2>&1; echo "pytest group rc=$?"; tail -3 $OUT/pytest_multi_n${N}_$TAG.log
MODES="peer nccl"; [[ " $* " == *" fuse "* ]] && MODES="peer nccl fused"
P=29650
for RED in $MODES; do
  for WL in cfg3 cfg4; do
    P=$((P+1))
    F=$OUT/bench_${WL}_n${N}_${RED}_$TAG
    case $RED in peer) FLAGS="--reduce auto";; nccl) FLAGS="--reduce nccl";; fused) FLAGS="--reduce auto --peer-fuse always";; esac
    timeout 300 $TR --master-port $P bench.py --gpus $N --workload $WL $FLAGS --no-supplied --no-reduced --no-cpu-baseline > $F.json 2> $F.err
    echo "bench $WL n$N $RED rc=$?"; python - "$F.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e ms", d["e2e"]["ms_per_eval"], "kernel ms", d["roofline"]["kernel_ms"], "|", d["config"]["sharding"], "|", d["clocks"])
except Exception as e:
    print("no line:", e)
PY
  done
done
if [[ " $* " == *" group "* ]]; then
  timeout 600 python tools/time_group.py > $OUT/group_n${N}_$TAG.jsonl 2> $OUT/group_n${N}_$TAG.err; echo "group rc=$?"; cat $OUT/group_n${N}_$TAG.jsonl
fi
