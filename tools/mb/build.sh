#!/bin/bash
# micro-benchmarks (not part of the library); binaries land in tools/mb/bin (git-ignored, shipped by gpurun)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/bin"
for f in "$HERE"/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o "$HERE/bin/$(basename "${f%.cu}")" "$f"
done
