#!/bin/bash
# Build libgrakel_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
set -o pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libgrakel_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -Xptxas -v \
  -o "$OUT" "$HERE/api.cu" -lcudart -lpthread -ldl 2>&1 | grep -v "^$" > "$HERE/../build.log" || { cat "$HERE/../build.log"; exit 1; }
echo "built $OUT"
# optional host-side accelerator: CPython-level packer for {(u, v): w} graph lists (packing.py falls back to numpy without it)
PYINC=$(python3 -c "import sysconfig; print(sysconfig.get_paths()['include'])")
PYEXT=$(python3 -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
if gcc -O2 -fPIC -shared -Wall -I"$PYINC" -o "$HERE/../_fastpack$PYEXT" "$HERE/fastpack.c" 2>> "$HERE/../build.log"; then
  echo "built $HERE/../_fastpack$PYEXT"
else
  echo "warning: _fastpack not built (see build.log); the numpy packer will be used"
fi
