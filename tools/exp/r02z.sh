#!/bin/bash
# new BK1 default (2 x 256 threads, 128 registers): GPU tests, sanitizer, volatile slot loads variant, bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log
tail -3 gpurun_out/r02z_pytest.log
L=gpurun_out/r02z_variants.log; : > $L
run() { m=$1; v=$2; n=8388608
  KINETIX_B200_TRUST_CACHE=1 timeout 300 python tools/quick_time.py --mech $m --n $n --reps 5 --cache build/variants/$v --tag "$m:$v" --check >> $L 2>&1; }
for v in swp wide wide_v; do run gri30 $v; done
grep -v "^$" $L | sed -E 's/\| BK2.*\| err/| err/' | cut -c1-160
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py gri30 NH3Konnov_edit H2_Konnov > gpurun_out/r02z_memcheck.txt 2>&1; tail -3 gpurun_out/r02z_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py gri30 > gpurun_out/r02z_racecheck.txt 2>&1; tail -3 gpurun_out/r02z_racecheck.txt
timeout 900 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02z_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','bk1_ms','bk2_ms')})
print('roofline', {k:d['roofline'].get(k) for k in ('kernel','frac','fp64_lane_instr_per_state','work_source')})
print('other', {k:d['roofline_other'].get(k) for k in ('kernel','frac','fp64_lane_instr_per_state','work_source')})
for c in d['configs']: print({k:(v if not isinstance(v,dict) else v.get('frac')) for k,v in c.items() if k in ('name','bk1_states_per_s','bk2_states_per_s','states_per_s','bk1_fp64','bk2_fp64')})
"
