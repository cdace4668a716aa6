#!/bin/bash
# more warps still: GRI-3.0 2 x 320 / 2 x 384 threads (96 / 80 registers); EtOHKonnov 384 threads x 168 registers with capped live sets
mkdir -p gpurun_out
L=gpurun_out/r02ab_variants.log; : > $L
run() { m=$1; v=$2; n=4194304
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in wide e320 e384; do run gri30 $v; done
for v in L2 x36s4 x40s4 x40s0 x44s4; do run EtOHKonnov $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
