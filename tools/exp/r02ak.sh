#!/bin/bash
# ptxas --register-usage-level on the GRI-3.0 module
mkdir -p gpurun_out
L=gpurun_out/r02ak_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in wide rul0 rul2 wide; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| thermo.*\| err/| err/' | cut -c1-200
