#!/bin/bash
# GPU call r02p: validation of the committed state (BK2 Wilke pipelining, heptane cold placement, EtOH two-halves BK2) + config-5 sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
tail -3 gpurun_out/r02p_pytest.log
timeout 600 ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:kx_bk --csv --log-file gpurun_out/r02p_counts.csv python tools/ncu_counts.py --run > gpurun_out/r02p_counts.log 2>&1
python tools/ncu_counts.py --parse gpurun_out/r02p_counts.csv gpurun_out/counts_run.json > gpurun_out/r02p_counts_parsed.log 2>&1
cp profiles/counts_r02.json gpurun_out/counts_r02.json
timeout 900 python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02p_bench_ref.json 2> gpurun_out/r02p_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02p_launches.csv \
    python bench.py --steps 2 --warmup 3 --n-states 4194304 --no-cpu-baseline --configs none > /dev/null 2>&1
python tools/launch_shares.py gpurun_out/r02p_launches.csv 'ncu --metrics gpu__time_duration.sum -c 400 python bench.py --steps 2 --warmup 3 --n-states 4194304 --configs none' > gpurun_out/r02p_launches_summary.txt
cat gpurun_out/r02p_launches_summary.txt
python -c "
import json
d=json.load(open('gpurun_out/r02p_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','bk1_ms','bk2_ms')})
print('roofline', {k:d['roofline'].get(k) for k in ('kernel','frac','fp64_lane_instr_per_state','traffic','algorithmic_bytes_per_launch','frac_reference_W')})
print('other', {k:d['roofline_other'].get(k) for k in ('kernel','frac','fp64_lane_instr_per_state','traffic','frac_reference_W')})
print('e2e', d['e2e']['value'], d['e2e']['two_call_value'])
for c in d['configs']: print({k:(v if not isinstance(v,dict) else v.get('frac')) for k,v in c.items() if k in ('name','bk1_states_per_s','bk2_states_per_s','states_per_s','bk1_fp64','bk2_fp64','bk1_hbm','bk1_mufu','thermo_ms','state_read_ms_upper_bound')})
print(d.get('cpu_baseline'))
"
# BASELINE config 5: precision / size sweep on one GPU (1 Ki .. 32 Mi states both kernels, 128 Mi states BK1)
timeout 900 python tools/sweep.py --mech gri30 --modes f64,fpmix,f32 --min 1024 --max 33554432 --out gpurun_out/r02p_sweep_gri.jsonl > /dev/null 2> gpurun_out/r02p_sweep.err
timeout 600 python tools/sweep.py --mech gri30 --modes f64,fpmix,f32 --min 134217728 --max 134217728 --kernels bk1 --out gpurun_out/r02p_sweep_gri128m.jsonl > /dev/null 2>> gpurun_out/r02p_sweep.err
python - <<'PY'
import json
for f in ('gpurun_out/r02p_sweep_gri.jsonl','gpurun_out/r02p_sweep_gri128m.jsonl'):
    for l in open(f):
        d=json.loads(l); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('mode','n_states','bk1_states_per_s','bk2_states_per_s','bk1_ms','bk2_ms')})
PY
