#!/bin/bash
# usage: build_variants.sh name "-DKSS=2 -DKSS_MINB=2" [name flags ...]  ->  scripts/gpu/variants/lib<name>.so
cd "$(dirname "$0")/../../fhe-si_b200/csrc" || exit 1
mkdir -p ../../scripts/gpu/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --expt-extended-lambda \
       -diag-suppress 550 -Xptxas -v $flags -o ../../scripts/gpu/variants/lib$name.so fhesi_lib.cu 2>&1 \
    | grep -A2 "k_fused_keyswitch_splitILb0\|k_fused_tensorILb0" | grep -v "^--" | sed "s/^/[$name] /" &
done
wait
