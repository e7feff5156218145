#!/bin/bash
# Build the experimental-kernel library (measurement aid, not shipped): scripts/exp/libevx_exp.so
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --threads 0 -shared \
  -Xcompiler -fPIC ${EXP_FLAGS} -o libevx_exp.so exp_*.cu
