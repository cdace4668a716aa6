#!/bin/bash
# GPU call r02e: BK2 column-loop unroll factors against the round-1 kernels on the same box
mkdir -p gpurun_out
run() { m=$1; v=$2; n=4194304; [ $m = gri30 ] && n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> gpurun_out/r02e_variants.log 2>&1; }
for v in cu1 cu3 cu9 r01; do run gri30 $v; done
for v in cu5 cu10 cu5_128 r01; do run EtOHKonnov $v; done
for v in cu4 cu8 r01; do run heptaneLu88 $v; done
grep -v "^$" gpurun_out/r02e_variants.log | cut -c1-220
