#!/bin/bash
# BK1 GRI-3.0: Arrhenius rate constants by a rolled, table-driven pre-pass (inline loop / real call) ahead of each chunk of units
mkdir -p gpurun_out
L=gpurun_out/r02ag_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in kf0 kf24u4 kf24u2 kc24u4 kc24u3 kc24u8 kc24u6 kc24u2 kc12u4; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
for v in kc24u4; do
timeout 300 ncu --set full --clock-control none -k regex:kx_bk1 -c 1 -o /tmp/full_$v python tools/quick_time.py --mech gri30 --n 4194304 --reps 1 --cache build/variants/$v > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_$v.ncu-rep > gpurun_out/r02ag_ncu_bk1_$v.txt 2>&1; cat gpurun_out/r02ag_ncu_bk1_$v.txt; done
