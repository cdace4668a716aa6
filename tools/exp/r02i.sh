#!/bin/bash
mkdir -p gpurun_out
run() { m=$1; v=$2; n=4194304; [ $m = gri30 ] && n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> gpurun_out/r02i_variants.log 2>&1; }
for v in g1 g3 r01; do run gri30 $v; done
for v in g5 g2 g5_128 r01; do run EtOHKonnov $v; done
for v in g8 g4 r01; do run heptaneLu88 $v; done
grep -v "^$" gpurun_out/r02i_variants.log | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
