#!/bin/bash
# build_variant.sh NAME [nvcc flags...] -> unfazed_b200/libunfazed_sm100_NAME.so (profiling variants; select with UNFZ_LIB)
set -e
cd "$(dirname "$0")/../unfazed_b200/csrc"
name=$1; shift
nvcc "$@" -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
  -o ../libunfazed_sm100_$name.so abi.cu scan.cu sites.cu reads.cu chain.cu insert.cu
