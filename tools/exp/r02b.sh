#!/bin/bash
# GPU call r02b: BK2 with the rolled column loop (A/B: prefetch on/off, CTA shapes), full GPU tests, bench, ncu capture
mkdir -p gpurun_out
for v in "" bk2_nopf bk2_st2 bk2_fullrcp; do
  c=""; [ -n "$v" ] && c="--cache build/variants/$v"
  timeout 300 python tools/quick_time.py --mech gri30 --n 8388608 --reps 5 $c --tag "gri30:${v:-default}" --check >> gpurun_out/r02b_variants.log 2>&1
done
for v in "" bk2_128 bk2_nopf; do
  c=""; [ -n "$v" ] && c="--cache build/variants/$v"
  timeout 300 python tools/quick_time.py --mech EtOHKonnov --n 4194304 --reps 3 $c --tag "etoh:${v:-default}" --check >> gpurun_out/r02b_variants.log 2>&1
done
for v in "" bk2_320 bk2_p2; do
  c=""; [ -n "$v" ] && c="--cache build/variants/$v"
  timeout 300 python tools/quick_time.py --mech heptaneLu88 --n 4194304 --reps 3 $c --tag "heptane:${v:-default}" --check >> gpurun_out/r02b_variants.log 2>&1
done
for m in NH3Konnov_edit gri30-35 gri30-27 chempolimi_edit; do
  timeout 300 python tools/quick_time.py --mech $m --n 4194304 --reps 3 --tag "$m" --check >> gpurun_out/r02b_variants.log 2>&1
done
grep -v "^$" gpurun_out/r02b_variants.log | cut -c1-220
timeout 400 ncu --set full --import-source on --clock-control none -k regex:kx_bk2 -c 1 -o gpurun_out/r02b_gri30_bk2 python tools/quick_time.py --mech gri30 --n 2097152 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02b_gri30_bk2.ncu-rep > gpurun_out/r02b_ncu_gri30_bk2.txt 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -8 gpurun_out/r02b_pytest.log
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r02b_bench.json
