#!/bin/sh
# Builds libphylo_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
exec "$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-fvisibility=hidden -shared -cudart static \
    -I../../include -o libphylo_b200.so kernels.cu phylo_b200.cu plan.cpp subst.cpp "$@"
