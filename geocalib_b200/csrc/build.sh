#!/bin/bash
# Builds libgclm.so for sm_100a in-tree (the .so is git-ignored but travels with gpurun).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${GCLM_OUT:-libgclm.so}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unknown-pragmas -shared -o "$OUT" gclm_engine.cu -ldl "$@"
