#!/bin/bash
# Build libgrakel_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libgrakel_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -Xptxas -v \
  -o "$OUT" "$HERE/api.cu" -lcudart -lpthread 2>&1 | grep -v "^$" > "$HERE/../build.log" || { cat "$HERE/../build.log"; exit 1; }
echo "built $OUT"
