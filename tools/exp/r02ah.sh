#!/bin/bash
# EtOHKonnov BK1: Arrhenius rate constants by the called, table-driven pre-pass into tensor-memory slots
mkdir -p gpurun_out
L=gpurun_out/r02ah_variants.log; : > $L
run() { m=$1; v=$2; n=4194304
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in L2 ek48u6 ek48u4 ek64u8; do run EtOHKonnov $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
for v in ek48u6; do
timeout 300 ncu --set full --clock-control none -k regex:kx_bk1 -c 1 -o /tmp/full_$v python tools/quick_time.py --mech EtOHKonnov --n 1048576 --reps 1 --cache build/variants/$v > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_$v.ncu-rep > gpurun_out/r02ah_ncu_bk1_$v.txt 2>&1; cat gpurun_out/r02ah_ncu_bk1_$v.txt; done
