#!/bin/sh
# Builds libphylo_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  kernels.cu is compiled in five
# parts in parallel (-DPHYLO_PART=0..4: each instantiates its share of the sweep kernel, see the note in kernels.cu);
# PHYLO_B200_SERIAL_BUILD=1 compiles it as one translation unit instead.  Extra arguments go to every nvcc call.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -I../../include"
if [ -n "$PHYLO_B200_SERIAL_BUILD" ]; then
    exec "$NVCC" $FLAGS -shared -cudart static -o libphylo_b200.so kernels.cu phylo_b200.cu plan.cpp subst.cpp "$@"
fi
OBJ=build
mkdir -p "$OBJ"
pids=""
for p in 0 1 2 3 4; do
    "$NVCC" $FLAGS -DPHYLO_PART=$p -c -o "$OBJ/kernels_$p.o" kernels.cu "$@" &
    pids="$pids $!"
done
"$NVCC" $FLAGS -c -o "$OBJ/phylo_b200.o" phylo_b200.cu "$@" &
pids="$pids $!"
"$NVCC" $FLAGS -c -o "$OBJ/plan.o" plan.cpp "$@" &
pids="$pids $!"
"$NVCC" $FLAGS -c -o "$OBJ/subst.o" subst.cpp "$@" &
pids="$pids $!"
rc=0
for pid in $pids; do wait "$pid" || rc=1; done
[ $rc -eq 0 ] || { echo "build.sh: a compile step failed" >&2; exit 1; }
exec "$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC -o libphylo_b200.so \
    "$OBJ/kernels_0.o" "$OBJ/kernels_1.o" "$OBJ/kernels_2.o" "$OBJ/kernels_3.o" "$OBJ/kernels_4.o" \
    "$OBJ/phylo_b200.o" "$OBJ/plan.o" "$OBJ/subst.o"
