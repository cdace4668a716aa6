#!/bin/bash
# BK2: conductivity sums moved from the first Wilke pass (serial head) to the second (independent dot products)
mkdir -p gpurun_out
L=gpurun_out/r02aq_variants.log; : > $L
run() { m=$1; v=$2; n=4194304; [ $m = gri30 ] && n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
run gri30 wide; run gri30 cnd; run gri30 wide; run gri30 cnd; run EtOHKonnov cnd; run heptaneLu88 cur; run heptaneLu88 cnd
grep -v "^$" $L | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
