#!/bin/bash
# Builds the product library (sm_100a only), the CPU oracle and the synthetic-input generator, in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRC="skani_b200/csrc/seeding.cu skani_b200/csrc/api.cu"
for f in skani_b200/csrc/screen.cu skani_b200/csrc/chain.cu skani_b200/csrc/todo_stubs.cu; do [ -f $f ] && SRC="$SRC $f"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 -shared \
  -o skani_b200/libskani_b200.so $SRC -lcudart
make -s -C oracle
/usr/bin/g++ -O3 -march=x86-64-v3 -std=c++17 -fPIC -fopenmp -shared -o bench_support/libsynth.so bench_support/synth.cpp
echo build ok
