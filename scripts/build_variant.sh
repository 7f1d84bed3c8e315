#!/bin/bash
# Builds a tuning variant of the library into scratch/variants/<name>.so (travels to the GPU box,
# git-ignored).  usage: scripts/build_variant.sh <name> <nvcc -D flags...>
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
cp $root/enrgy_b200/csrc/*.cu $root/enrgy_b200/csrc/*.cuh $tmp/
mkdir -p $tmp/../../include 2>/dev/null || true
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -I$root/include"
mkdir -p $root/scratch/variants
# sources include "../../include/enrgy_b200.h": compile in place of the tree layout
mkdir -p $tmp/a/b && cp $tmp/*.cu $tmp/*.cuh $tmp/a/b/ && mkdir -p $tmp/include && cp $root/include/enrgy_b200.h $tmp/include/
cd $tmp/a/b
for f in kernels shade prepass api; do $NVCC $FLAGS "$@" ${PTXAS_V:+-Xptxas -v} -c $f.cu -o $f.o & done
wait
$NVCC -shared -o $root/scratch/variants/$name.so kernels.o shade.o prepass.o api.o -cudart static
rm -rf $tmp
echo built scratch/variants/$name.so
