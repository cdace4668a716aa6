#!/bin/bash
# BK1 GRI-3.0 2 x 256 x 128 registers: wdot_k of rarely used species in slots as well
mkdir -p gpurun_out
L=gpurun_out/r02w_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in swp w0 w5 w10 w20 w40; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
