#!/bin/bash
# Builds libqcart.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libqcart.so"
OBJ="$HERE/build"
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
pids=()
for f in qc_kernels.cu qc_api.cu qc_model.cpp; do
  $NVCC $FLAGS ${QC_PTXAS_V:+-Xptxas -v} -c -o "$OBJ/${f%.*}.o" "$HERE/$f" > "$OBJ/${f%.*}.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
cat "$OBJ"/*.log
[ $rc -eq 0 ] || { echo "compile failed"; exit 1; }
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ/qc_kernels.o" "$OBJ/qc_api.o" "$OBJ/qc_model.o" -cudart static
echo "built $OUT"
