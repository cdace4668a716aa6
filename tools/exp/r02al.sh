#!/bin/bash
# BK1 GRI-3.0 wide layout with a capped live set
mkdir -p gpurun_out
L=gpurun_out/r02al_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in lc0 lc18 lc21 lc23; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
