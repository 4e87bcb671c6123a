#!/bin/bash
for d in build/ab/*/; do n=$(basename $d); echo "== $n"; NLCUDA_LIB=$PWD/$d/libnlcuda.so python scripts/exp_slab_local.py; done
