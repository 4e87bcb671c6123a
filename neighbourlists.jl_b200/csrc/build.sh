#!/bin/bash
# Builds libnlcuda.so for sm_100a in-tree (next to the package's __init__.py).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libnlcuda.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
CCBIN=""
if [ -x /usr/bin/g++ ]; then CCBIN="-ccbin /usr/bin/g++"; fi
CXX="${CXX:-g++}"
# host side of the device -> host transfer format: plain C++, AVX2 variants behind target attributes
"$CXX" -O3 -std=c++17 -fPIC -fvisibility=hidden -c "$HERE/nl_hostcodec.cpp" -o "$HERE/nl_hostcodec.o"
"$NVCC" $CCBIN -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared ${NL_NVCC_EXTRA:-} \
  -o "$OUT" "$HERE/nlcuda.cu" "$HERE/nl_hostcodec.o"
rm -f "$HERE/nl_hostcodec.o"
echo "built $OUT"
