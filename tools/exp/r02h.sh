#!/bin/bash
mkdir -p gpurun_out
for v in ka1 ka2 ka4; do
  timeout 300 python tools/quick_time.py --mech gri30 --n 8388608 --reps 5 --cache build/variants/$v --tag "gri30:$v" --check >> gpurun_out/r02h_variants.log 2>&1
done
grep -v "^$" gpurun_out/r02h_variants.log | cut -c1-120
