#!/bin/bash
# A/B of library variants in build/ab: step / count / fill times at the headline size (3 runs each, interleaved)
for rep in 1 2 3; do
for d in build/ab/*/; do
  n=$(basename $d)
  echo -n "$n: "; NLCUDA_LIB=$PWD/$d/libnlcuda.so python scripts/exp_fill.py 2>/dev/null | head -1
done
done
