#!/bin/bash
mkdir -p gpurun_out
run() { m=$1; v=$2; n=4194304; [ $m = gri30 ] && n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> gpurun_out/r02j_variants.log 2>&1; }
for v in L2 L1_128 r01; do run EtOHKonnov $v; done
run gri30 cur; run heptaneLu88 cur
grep -v "^$" gpurun_out/r02j_variants.log | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_bk2 -c 1 -o /tmp/full_etoh_bk2 python tools/quick_time.py --mech EtOHKonnov --n 524288 --reps 1 --cache build/variants/L2 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_etoh_bk2.ncu-rep > gpurun_out/r02j_ncu_etoh_bk2_L2.txt 2>&1
cat gpurun_out/r02j_ncu_etoh_bk2_L2.txt
