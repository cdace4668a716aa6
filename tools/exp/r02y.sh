#!/bin/bash
# the 2 x 256-thread / 128-register BK1 layout on the small mechanisms
mkdir -p gpurun_out
L=gpurun_out/r02y_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for m in gri30-20 H2_Konnov H2_new_mech LiDryer; do for v in cur d2ct0; do run $m $v; done; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
