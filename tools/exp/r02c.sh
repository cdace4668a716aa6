#!/bin/bash
# GPU call r02c: BK2 with 16-byte coefficient loads; exp-polynomial shapes in BK1; tests; bench
mkdir -p gpurun_out
for v in "" bk2_st2 evenodd estrin; do
  c=""; [ -n "$v" ] && c="--cache build/variants/$v"
  timeout 300 python tools/quick_time.py --mech gri30 --n 8388608 --reps 5 $c --tag "gri30:${v:-default}" --check >> gpurun_out/r02c_variants.log 2>&1
done
for v in "" bk2_128 evenodd sync2; do
  c=""; [ -n "$v" ] && c="--cache build/variants/$v"
  timeout 300 python tools/quick_time.py --mech EtOHKonnov --n 4194304 --reps 3 $c --tag "etoh:${v:-default}" --check >> gpurun_out/r02c_variants.log 2>&1
done
for v in "" evenodd; do
  c=""; [ -n "$v" ] && c="--cache build/variants/$v"
  timeout 300 python tools/quick_time.py --mech heptaneLu88 --n 4194304 --reps 3 $c --tag "heptane:${v:-default}" --check >> gpurun_out/r02c_variants.log 2>&1
done
grep -v "^$" gpurun_out/r02c_variants.log | cut -c1-220
timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_bk2 -c 1 -o gpurun_out/r02c_gri30_bk2 python tools/quick_time.py --mech gri30 --n 2097152 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02c_gri30_bk2.ncu-rep > gpurun_out/r02c_ncu_gri30_bk2.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -8 gpurun_out/r02c_pytest.log
