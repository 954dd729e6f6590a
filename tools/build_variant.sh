#!/bin/bash
# Builds a kernel variant: tools/build_variant.sh <name> <extra nvcc -D flags...>  ->  build/librcvvote_<name>.so
n=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared "$@" -o build/librcvvote_$n.so rcvpose_b200/csrc/*.cu
