#!/bin/bash
# BK1 GRI-3.0: four 128-thread CTAs per SM at 128 registers (4 warps per scheduler), slots in shared + tensor memory, C_k in slots
mkdir -p gpurun_out
L=gpurun_out/r02s_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in swp q4 q4c q4ct q4c30; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
for v in q4ct; do
timeout 300 ncu --set full --clock-control none -k regex:kx_bk1 -c 1 -o /tmp/full_$v python tools/quick_time.py --mech gri30 --n 4194304 --reps 1 --cache build/variants/$v > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_$v.ncu-rep > gpurun_out/r02s_ncu_bk1_$v.txt 2>&1; cat gpurun_out/r02s_ncu_bk1_$v.txt; done
