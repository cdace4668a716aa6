#!/bin/bash
# BK1 GRI-3.0: neighbourhood of the 2 x 256-thread / 128-register layout
mkdir -p gpurun_out
L=gpurun_out/r02u_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in swp d2ct16 d2ct24 d2ct32 d2ct12 d2c16 d2ct16p2 d2ct16p6 d1ct16 d2ct0; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
for v in d2ct16; do
timeout 300 ncu --set full --clock-control none -k regex:kx_bk1 -c 1 -o /tmp/full_$v python tools/quick_time.py --mech gri30 --n 4194304 --reps 1 --cache build/variants/$v > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_$v.ncu-rep > gpurun_out/r02u_ncu_bk1_$v.txt 2>&1; cat gpurun_out/r02u_ncu_bk1_$v.txt; done
