#!/bin/bash
# Builds libqcart.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libqcart.so"
OBJ="$HERE/build"
# QC_DEBUG_HOOKS=1: development build with the QCART_DEBUG timing hooks (wrong results by design) -> libqcart_dbg.so, never loaded by the package
if [ -n "$QC_DEBUG_HOOKS" ]; then OUT="$HERE/../libqcart_dbg.so"; OBJ="$HERE/build_dbg"; fi
# QC_VARIANT=name QC_DEFS="-DX=1 ...": experimental build with extra defines -> libqcart_name.so (A/B tests through QCART_LIB)
if [ -n "$QC_VARIANT" ]; then OUT="$HERE/../libqcart_$QC_VARIANT.so"; OBJ="$HERE/build_$QC_VARIANT"; fi
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${QC_DEBUG_HOOKS:+-DQC_DEBUG_HOOKS} $QC_DEFS"
pids=()
SRCS="qc_kernels.cu qc_api.cu qc_rollout.cu qc_model.cpp $(cd "$HERE" && ls qc_inst_*.cu)"
for f in $SRCS; do
  $NVCC $FLAGS ${QC_PTXAS_V:+-Xptxas -v} -c -o "$OBJ/${f%.*}.o" "$HERE/$f" > "$OBJ/${f%.*}.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
for f in $SRCS; do cat "$OBJ/${f%.*}.log"; done
[ $rc -eq 0 ] || { echo "compile failed"; exit 1; }
OBJS=""; for f in $SRCS; do OBJS="$OBJS $OBJ/${f%.*}.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $OBJS -cudart static
echo "built $OUT"
