#!/bin/bash
# Builds libenrgy_b200.so in-tree for sm_100a.  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wno-unused-function"
# the four translation units compile side by side (kernels.cu alone takes most of the time)
pids=()
$NVCC $FLAGS $EXTRA ${PTXAS_V:+-Xptxas -v} -c kernels.cu -o kernels.o & pids+=($!)
$NVCC $FLAGS $EXTRA ${PTXAS_V:+-Xptxas -v} -c shade.cu -o shade.o & pids+=($!)
$NVCC $FLAGS $EXTRA -c prepass.cu -o prepass.o & pids+=($!)
$NVCC $FLAGS $EXTRA -c api.cu -o api.o & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
OUT=${OUT:-libenrgy_b200.so}
$NVCC -shared -o $OUT kernels.o shade.o prepass.o api.o -cudart static
echo built $(pwd)/$OUT
