#!/bin/bash
# BK2 Wilke passes software-pipelined (branch-free species head, three dot products in the second pass)
mkdir -p gpurun_out
L=gpurun_out/r02l_variants.log; : > $L
run() { m=$1; v=$2; n=4194304; [ $m = gri30 ] && n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
run gri30 noswp; run gri30 swp; run gri30 noswp; run gri30 swp
run EtOHKonnov L2; run EtOHKonnov swp
run heptaneLu88 cur; run heptaneLu88 swp
grep -v "^$" $L | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_bk2 -c 1 -o /tmp/full_gri_bk2 python tools/quick_time.py --mech gri30 --n 1048576 --reps 1 --cache build/variants/swp > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_gri_bk2.ncu-rep > gpurun_out/r02l_ncu_gri_bk2_swp.txt 2>&1
cat gpurun_out/r02l_ncu_gri_bk2_swp.txt
