#!/bin/bash
# Build a variant of libxrt_b200.so HERE (nvcc cross-compiles) with extra -D flags for ONE source file:
#   tools/build_variant.sh TAG reflect.cu "-DXRTB_REFLECT_MINBLOCKS=3"
# -> variants/TAG/libxrt_b200.so (travels to the GPU box; load it with XRT_B200_LIB=variants/TAG/libxrt_b200.so).
# The other objects come from build/obj (run __graft_entry__.build() first).
set -e
TAG=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/variants/$TAG"
FMAD=""
case "$SRC" in reflect.cu|rays.cu|source.cu|undulator.cu|undulator_ext.cu) FMAD="-fmad=false";; esac
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 $FMAD "$@" \
  -Xptxas -v -I "$ROOT/include" -c -o "$ROOT/variants/$TAG/${SRC%.cu}.o" "$ROOT/xrt_b200/csrc/$SRC" 2> "$ROOT/variants/$TAG/ptxas.log"
OBJS=""
for o in "$ROOT"/build/obj/*.o; do
  if [ "$(basename $o)" != "${SRC%.cu}.o" ]; then OBJS="$OBJS $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/variants/$TAG/libxrt_b200.so" "$ROOT/variants/$TAG/${SRC%.cu}.o" $OBJS -ldl
rm -f "$ROOT/variants/$TAG/${SRC%.cu}.o"
echo "built variants/$TAG/libxrt_b200.so"
