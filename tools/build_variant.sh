#!/bin/bash
# Build a tuning variant of libsaa_b200.so: tools/build_variant.sh NAME -DSAA_COPY=3 ...  ->  build/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC,-fopenmp -lgomp \
  "$@" riskaversetrajopt_b200/csrc/saa_b200.cu -o build/$name.so 2>&1 | grep -v "^$" | grep -i "error\|registers" || true
