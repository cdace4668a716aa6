#!/bin/bash
# BK2: the next row block's column sums fetched from tensor memory during the diagonal tile / store section
mkdir -p gpurun_out
L=gpurun_out/r02ar_variants.log; : > $L
run() { m=$1; v=$2; n=4194304; [ $m = gri30 ] && n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
run gri30 wide; run gri30 tmpf; run gri30 wide; run gri30 tmpf; run EtOHKonnov tmpf; run heptaneLu88 cur; run heptaneLu88 tmpf
grep -v "^$" $L | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
