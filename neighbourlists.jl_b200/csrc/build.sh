#!/bin/bash
# Builds libnlcuda.so for sm_100a in-tree (next to the package's __init__.py).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libnlcuda.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
CCBIN=""
if [ -x /usr/bin/g++ ]; then CCBIN="-ccbin /usr/bin/g++"; fi
"$NVCC" $CCBIN -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared ${NL_NVCC_EXTRA:-} \
  -o "$OUT" "$HERE/nlcuda.cu"
echo "built $OUT"
