#!/bin/bash
# One GPU visit: parity tests, smoke, bench (both arms; the default line carries cfg1/cfg2/cfg4 as extra legs), ncu launch
# list, ncu full captures of the two rollout kernels at the bench's sizes.   Usage (through gpurun): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"; cat $OUT/bench_ref_$TAG.json
if [ "${2:-}" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rollout_philox -s 3 -c 1 -f -o $OUT/prof_philox_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-supplied --no-reduced --no-extra > $OUT/ncu_full_philox_$TAG.log 2>&1; echo "ncu philox rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rollout_supplied_bulk -s 3 -c 1 -f -o $OUT/prof_supplied_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reduced --no-extra > $OUT/ncu_full_supplied_$TAG.log 2>&1; echo "ncu supplied rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rollout_aug_thread -s 20 -c 1 -f -o $OUT/prof_aug_thread_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-supplied > $OUT/ncu_full_aug_$TAG.log 2>&1; echo "ncu aug thread (N = 4) rc=$?"
fi
ls -la $OUT | tail -12
