#!/bin/bash
# runs bench.py once per variant in build/ab and prints ms/step + stage times
for d in build/ab/*/; do
  n=$(basename $d)
  NLCUDA_LIB=$PWD/$d/libnlcuda.so python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$n', round(d['ms_per_step'],3), d['config']['stage_ms'])"
done
