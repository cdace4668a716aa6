#!/bin/bash
# BK1 GRI-3.0: two 256-thread CTAs per SM at 128 registers (4 warps per scheduler, two instruction streams)
mkdir -p gpurun_out
L=gpurun_out/r02t_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in swp d2ct16 d2ct8 d2c30_16 d2c30_32; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
