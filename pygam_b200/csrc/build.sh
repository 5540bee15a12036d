#!/bin/bash
# Build the C-ABI engine for sm_100a, in-tree: libpgb.so (the product) and libpgb_selfcheck.so (the same objects
# plus the host-side self check of the cell plan; loaded by the CPU test-suite only).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 20013,20015 $PGB_NVCC_EXTRA"
mkdir -p build
pids=()
for f in pgb_core pgb_pass pgb_pass_b pgb_gram pgb_cells pgb_solve pgb_upload pgb_selfcheck; do
  $NVCC $FLAGS -c $f.cu -o build/$f.o "$@" > build/$f.log 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
if [ $rc -ne 0 ]; then cat build/*.log; exit 1; fi
$NVCC -shared -o libpgb.so build/pgb_core.o build/pgb_pass.o build/pgb_pass_b.o build/pgb_gram.o build/pgb_cells.o build/pgb_solve.o build/pgb_upload.o
$NVCC -shared -o libpgb_selfcheck.so build/pgb_core.o build/pgb_pass.o build/pgb_pass_b.o build/pgb_gram.o build/pgb_cells.o build/pgb_solve.o build/pgb_upload.o build/pgb_selfcheck.o
echo "built $(pwd)/libpgb.so and libpgb_selfcheck.so"
