#!/bin/bash
# FP32-math BK1 (fpmix) GRI-3.0: CTA shapes
mkdir -p gpurun_out
L=gpurun_out/r02ai_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --sp 0 >> $L 2>&1; }
for v in sp_cur sp_256x2 sp_192x2 sp_256x1 sp_384x1 sp_128x4 sp_cur; do run gri30 $v; done
grep -v "^$" $L | cut -c1-120
