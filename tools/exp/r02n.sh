#!/bin/bash
# timing probe: BK1 body without the per-species NASA polynomial / exp (results are wrong by construction)
mkdir -p gpurun_out
L=gpurun_out/r02n_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" >> $L 2>&1; }
run gri30 swp; run gri30 probe_ng; run gri30 swp; run gri30 probe_ng
grep -v "^$" $L | cut -c1-120
