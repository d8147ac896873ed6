#!/bin/bash
# build_variants.sh name:flags ...  -> build_variants/libkfbo_<name>.so (kernel tuning experiments; timed by tools/variant_round.sh)
cd "$(dirname "$0")/../kf_bayesopt_b200/csrc" || exit 1
mkdir -p ../../build_variants
for v in "$@"; do
    n=${v%%:*}; f=${v#*:}
    ( nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $f -shared \
        -o ../../build_variants/libkfbo_$n.so kfbo_api.cu kfbo_kernels.cu -ldl 2> /tmp/build_$n.log || { echo "$n: BUILD FAILED"; tail -5 /tmp/build_$n.log; }
      echo "$n: $(grep -A3 'Compiling entry function .*rollout_philoxILb1' /tmp/build_$n.log | grep -E 'Used|spill' | tr '\n' ' ')" ) &
done
wait
