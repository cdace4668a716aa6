#!/bin/bash
# BK2 GRI-3.0: 512 threads as two halves sharing 256 one-state threads' states (four warps per scheduler at 128 registers)
mkdir -p gpurun_out
L=gpurun_out/r02ad_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in swp l2t10 l2t8 l2t6; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
for v in l2t8; do
timeout 300 ncu --set full --clock-control none -k regex:kx_bk2 -c 1 -o /tmp/full_$v python tools/quick_time.py --mech gri30 --n 4194304 --reps 1 --cache build/variants/$v > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_$v.ncu-rep > gpurun_out/r02ad_ncu_bk2_$v.txt 2>&1; cat gpurun_out/r02ad_ncu_bk2_$v.txt; done
