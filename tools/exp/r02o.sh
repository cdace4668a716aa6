#!/bin/bash
# BK1 GRI-3.0: CTA shape x barrier interval (more warps on one instruction window)
mkdir -p gpurun_out
L=gpurun_out/r02o_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in swp b384s4 b384s8 b384s16 b192s8 b192s4 b256s8; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
