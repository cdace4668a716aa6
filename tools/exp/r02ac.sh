#!/bin/bash
# small-launch BK1 instantiation: GPU tests, sweep of the small sizes, sanitizer on both kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_pytest.log
tail -3 gpurun_out/r02ac_pytest.log; grep "at_the_switch\|one wave" gpurun_out/r02ac_pytest.log | head -20
timeout 900 python tools/sweep.py --mech gri30 --modes f64 --min 1024 --max 4194304 --out gpurun_out/r02ac_sweep_gri.jsonl > /dev/null 2> gpurun_out/r02ac_sweep.err
python -c "
import json
for l in open('gpurun_out/r02ac_sweep_gri.jsonl'):
    d=json.loads(l); print(d['n_states_per_gpu'], round(d['bk1_ms']*1e3,1),'us', round(d['bk1_states_per_s']/1e6,1), 'M st/s | bk2', round(d['bk2_ms']*1e3,1),'us')
"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py gri30 > gpurun_out/r02ac_memcheck.txt 2>&1; tail -2 gpurun_out/r02ac_memcheck.txt
