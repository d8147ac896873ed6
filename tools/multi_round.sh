#!/bin/bash
# Multi-GPU visit (round 2): the multi-GPU tests, tools/multi_gpu_check.py, the driver's bench command at N (its line carries
# parity_n, the cfg4 leg and the NCCL-path timing), and the same with graph replay / host polling switched off one at a time.     Usage (through gpurun --gpus N): bash tools/multi_round.sh <N> <tag>
set -u
N=${1:-2}; TAG=${2:-r2}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/pytest_multi_n${N}_$TAG.log 2>&1; echo "pytest multi rc=$?"; tail -3 $OUT/pytest_multi_n${N}_$TAG.log
timeout 600 $TR --master-port 29660 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench n$N rc=$?"
python - $OUT/bench_n${N}_$TAG.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "nccl_ms")}, "e2e ms", d["e2e"]["ms_per_eval"], "kernel ms", d["roofline"]["kernel_ms"], "|", d["config"]["sharding"])
    print("parity_n", d.get("parity_n")); print("cfg4", d["extra"]["cfg4"]); print("host", d["host_phases_us"], d["clocks"])
except Exception as e:
    print("no line:", e, open(sys.argv[1]).read()[-2000:])
PY
timeout 300 $TR --master-port 29665 tools/soak.py ${SOAK_SECONDS:-15} 2>&1 | grep -E "soak ok|Error|error|assert" | tee $OUT/soak_n${N}_$TAG.txt; echo "soak n$N rc=${PIPESTATUS[0]}"
P=29670
for V in "graph_off KFBO_BENCH_GRAPH=0" "poll_off KFBO_BENCH_POLL=0"; do
  set -- $V; P=$((P+1))
  env $2 timeout 300 $TR --master-port $P bench.py --gpus $N --steps 20 --warmup 3 --no-extra --no-supplied --no-cpu-baseline > $OUT/bench_n${N}_${1}_$TAG.json 2> $OUT/bench_n${N}_${1}_$TAG.err
  echo "bench n$N $1 rc=$?"; python - $OUT/bench_n${N}_${1}_$TAG.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step")}, "e2e ms", d["e2e"]["ms_per_eval"], "kernel ms", d["roofline"]["kernel_ms"], d["host_phases_us"])
except Exception as e:
    print("no line:", e)
PY
done
