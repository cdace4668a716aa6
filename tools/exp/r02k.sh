#!/bin/bash
# conc-only cold placement (C_k in a slot, wdot_k in a register) across EtOHKonnov / GRI-3.0 / heptaneLu88
mkdir -p gpurun_out
L=gpurun_out/r02k_variants.log; : > $L
run() { m=$1; v=$2; n=4194304; [ $m = gri30 ] && n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in L2 c40 c40t call_t c40t_l70 c80t_l80 c20t call; do run EtOHKonnov $v; done
for v in cur g_c10 g_c20 g_c40 g_c400; do run gri30 $v; done
for v in cur h_c40c h_c400c h_c400c84; do run heptaneLu88 $v; done
grep -v "^$" $L | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
