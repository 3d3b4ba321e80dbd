#!/bin/bash
# N Ci D H W Co k pad [mode]
for c in "1 16 16 16 16 16 3 1" "1 16 16 16 16 8 3 1" "1 16 16 16 16 16 3 1 fwd" "1 16 16 16 16 8 3 1 fwd"; do
  for h in 1 0; do
    CFUN_TC_HALO=$h timeout 120 python tools/wg_case.py $c 2>&1 | grep -E "CASE \[" | sed "s/^/halo=$h /" | head -2
  done
done
