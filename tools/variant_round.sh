#!/bin/bash
# Time kernel build variants (build_variants/libkfbo_<name>.so, selected through KFBO_LIB_PATH) on the cfg3 workload,
# run the register-port microbenchmark, then the GPU parity tests against one chosen variant.
#   Usage (through gpurun): bash tools/variant_round.sh <tag> <variant to test> <variants to time...>
set -u
TAG=$1; TESTV=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for v in "$@" "$@"; do
    KFBO_LIB_PATH=$PWD/build_variants/libkfbo_$v.so python tools/time_eval.py 2>&1 | tail -1
done | tee $OUT/variants_$TAG.txt
if [ -x build_variants/rf_ports ]; then ./build_variants/rf_ports | tee $OUT/rf_ports_$TAG.txt; fi
if [ "$TESTV" != "none" ]; then
    KFBO_LIB_PATH=$PWD/build_variants/libkfbo_$TESTV.so timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest($TESTV) rc=$?"
    tail -5 $OUT/pytest_gpu_$TAG.log
fi
