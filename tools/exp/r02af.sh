#!/bin/bash
# final evidence of the round: GPU tests, smoke, executed-work table, bench line, reference arm, launch list, full captures, sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02af_pytest.log; tail -2 gpurun_out/r02af_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02af_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02af_smoke.log
timeout 600 ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:kx_bk --csv --log-file gpurun_out/r02af_counts.csv python tools/ncu_counts.py --run > gpurun_out/r02af_counts.log 2>&1
python tools/ncu_counts.py --parse gpurun_out/r02af_counts.csv gpurun_out/counts_run.json > gpurun_out/r02af_counts_parsed.log 2>&1
cp profiles/counts_r02.json gpurun_out/counts_r02.json
timeout 900 python bench.py > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02af_bench_ref.json 2> gpurun_out/r02af_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02af_launches.csv \
    python bench.py --steps 2 --warmup 3 --n-states 4194304 --no-cpu-baseline --configs none > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/r02af_launches.csv 'ncu --metrics gpu__time_duration.sum -c 400 python bench.py --steps 2 --warmup 3 --n-states 4194304 --configs none' > gpurun_out/r02af_launches_summary.txt
cat gpurun_out/r02af_launches_summary.txt | head -5
for k in bk1 bk2; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:kx_$k -c 1 -o /tmp/full_$k python tools/quick_time.py --mech gri30 --n 4194304 --reps 1 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/full_$k.ncu-rep > gpurun_out/r02af_ncu_full_$k.txt 2>&1; cat gpurun_out/r02af_ncu_full_$k.txt
done
python tools/ncu_stalls.py /tmp/full_bk1.ncu-rep > gpurun_out/r02af_stalls_bk1.txt 2>&1
timeout 900 python tools/sweep.py --mech gri30 --modes f64,fpmix,f32 --min 1024 --max 33554432 --out gpurun_out/r02af_sweep_gri.jsonl > /dev/null 2> gpurun_out/r02af_sweep.err
timeout 600 python tools/sweep.py --mech gri30 --modes f64,fpmix,f32 --min 134217728 --max 134217728 --kernels bk1 --out gpurun_out/r02af_sweep_gri128m.jsonl > /dev/null 2>> gpurun_out/r02af_sweep.err
python -c "
import json
d=json.load(open('gpurun_out/r02af_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','bk1_ms','bk2_ms')})
print('roofline', {k:d['roofline'].get(k) for k in ('kernel','frac','fp64_lane_instr_per_state','traffic','work_source')})
print('other', {k:d['roofline_other'].get(k) for k in ('kernel','frac','fp64_lane_instr_per_state','traffic','work_source')})
print('e2e', d['e2e']['value'], d['e2e']['two_call_value']); print(d['clocks'])
"
