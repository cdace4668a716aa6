#!/bin/bash
# BK1 GRI-3.0: the slot layout at 2 x 192 threads / 168 registers and 2 x 224 threads vs 2 x 256 / 128
mkdir -p gpurun_out
L=gpurun_out/r02ae_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in t256 t224 t192 t256; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
