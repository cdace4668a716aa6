#!/bin/bash
# GPU call r02f: GRI-3.0 BK1 with cold slots (spills removed), full GPU tests, bench, executed-work table
mkdir -p gpurun_out
for v in "" cold20 cold40; do
  c=""; [ -n "$v" ] && c="--cache build/variants/$v"
  timeout 300 python tools/quick_time.py --mech gri30 --n 8388608 --reps 5 $c --tag "gri30:${v:-default}" --check >> gpurun_out/r02f_variants.log 2>&1
done
grep -v "^$" gpurun_out/r02f_variants.log | cut -c1-220
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -4 gpurun_out/r02f_pytest.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_bk1 -c 1 -o /tmp/full_bk1_cold python tools/quick_time.py --mech gri30 --n 2097152 --reps 1 --cache build/variants/cold20 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_bk1_cold.ncu-rep > gpurun_out/r02f_ncu_gri30_bk1_cold20.txt 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_bk1 -c 1 -o /tmp/full_bk1 python tools/quick_time.py --mech gri30 --n 2097152 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/full_bk1.ncu-rep > gpurun_out/r02f_ncu_gri30_bk1.txt 2>&1
cat gpurun_out/r02f_ncu_gri30_bk1_cold20.txt gpurun_out/r02f_ncu_gri30_bk1.txt | grep -E "duration|fp64 pipe|dram|icache|stalls|local"
