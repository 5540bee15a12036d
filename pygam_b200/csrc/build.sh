#!/bin/bash
# Build libpgb.so (the C-ABI engine) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
$NVCC $FLAGS -shared -o libpgb.so pgb_core.cu pgb_gram.cu pgb_cells.cu pgb_solve.cu "$@"
echo "built $(pwd)/libpgb.so"
