#!/bin/bash
# GPU call r02d: restructured BK2 (32-bit counters) against the round-1 kernels on the same box
mkdir -p gpurun_out
for m in gri30 EtOHKonnov heptaneLu88; do
  n=4194304; [ $m = gri30 ] && n=8388608
  timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --tag "$m:default" --check >> gpurun_out/r02d_variants.log 2>&1
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/r01 --tag "$m:r01" --check >> gpurun_out/r02d_variants.log 2>&1
done
timeout 300 python tools/quick_time.py --mech gri30 --n 8388608 --reps 5 --cache build/variants/bk2_st2 --tag "gri30:st2" --check >> gpurun_out/r02d_variants.log 2>&1
grep -v "^$" gpurun_out/r02d_variants.log | cut -c1-220
timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_bk2 -c 1 -o gpurun_out/r02d_gri30_bk2 python tools/quick_time.py --mech gri30 --n 2097152 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02d_gri30_bk2.ncu-rep > gpurun_out/r02d_ncu_gri30_bk2.txt 2>&1
KINETIX_B200_TRUST_CACHE=1 timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_bk2 -c 1 -o gpurun_out/r02d_gri30_bk2_r01 python tools/quick_time.py --mech gri30 --n 2097152 --reps 1 --cache build/variants/r01 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02d_gri30_bk2_r01.ncu-rep > gpurun_out/r02d_ncu_gri30_bk2_r01.txt 2>&1
cat gpurun_out/r02d_ncu_gri30_bk2.txt gpurun_out/r02d_ncu_gri30_bk2_r01.txt
