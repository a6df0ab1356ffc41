#!/bin/sh
# A/B builds of libphylo_b200.so with other -D switches (kernels.cu: PHYLO_RSM, PHYLO_PF, PHYLO_PRETIP):
#   build_variant.sh <name> [-D...]   ->  variants/libphylo_b200_<name>.so   (select with PHYLO_B200_LIB=...)
set -e
cd "$(dirname "$0")"
name=$1; shift
mkdir -p variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
exec "$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-fvisibility=hidden -shared -cudart static \
    -I../../include -o "variants/libphylo_b200_$name.so" kernels.cu phylo_b200.cu plan.cpp subst.cpp "$@"
