#!/bin/bash
# A/B helper: builds library variants with different -D flags into build/ab/<name>/libnlcuda.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  NL_NVCC_EXTRA="$flags" bash neighbourlists.jl_b200/csrc/build.sh > /dev/null
  mkdir -p "build/ab/$name"
  cp neighbourlists.jl_b200/libnlcuda.so "build/ab/$name/libnlcuda.so"
done
