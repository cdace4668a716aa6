#!/bin/bash
# BK2: the last, partial round of batches as a second launch of the one-state-per-thread instantiation
mkdir -p gpurun_out
L=gpurun_out/r02ap_variants.log; : > $L
run() { m=$1; v=$2; n=$3
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 20 --cache build/variants/$v --tag "$m:$v:$n" --check >> $L 2>&1; }
for n in 917504 1245184 8388608; do run gri30 wide $n; run gri30 tail $n; done
grep -v "^$" $L | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
