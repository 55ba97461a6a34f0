#!/bin/bash
# Builds blend-kernel variants ON the GPU box and prints their blend times (tuning aid, not a bench).
cd instantsplat_b200/csrc
for v in "256 7 6" "128 7 6" "128 9 8" "256 8 7" "512 6 6" "128 10 9"; do
  set -- $v
  make clean > /dev/null 2>&1
  make -j8 EXTRA="-DGSB_CHUNK=$1 -DGSB_FWD_MINB=$2 -DGSB_BWD_MINB=$3" > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  grep -A2 "k_blend_...2ILb1" gs_raster.ptxas.log | grep -E "Used|spill" | tr '\n' ' '
  (cd ../.. && timeout 200 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('VARIANT $v ->', round(d['value'],1), 'it/s fwd', k['blend_fwd']['ms'], 'bwd', k['blend_bwd']['ms'])")
done
make clean > /dev/null 2>&1; make -j8 > /dev/null 2>&1
