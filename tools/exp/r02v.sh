#!/bin/bash
# BK1 GRI-3.0 2 x 256 x 128 registers: uniform TMEM base, TMEM / shared-memory split by use count
mkdir -p gpurun_out
L=gpurun_out/r02v_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in swp d2ct0 d2ct0u u10 u20 u30 u50; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
