#!/bin/bash
# BK2 two-halves kernel: which half a warp is decided once per diagonal tile / store block, not per pair
mkdir -p gpurun_out
L=gpurun_out/r02an_variants.log; : > $L
run() { m=$1; v=$2; n=4194304; [ $m = gri30 ] && n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
run EtOHKonnov dg2; run EtOHKonnov dg3; run EtOHKonnov dg3
grep -v "^$" $L | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
