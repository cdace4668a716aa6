#!/bin/bash
# 2-GPU call: the tests that need two devices, bench at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02_n2_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "two_devices or worker_processes" > gpurun_out/r02_n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_n2_pytest.log
tail -5 gpurun_out/r02_n2_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_n2_bench.json 2> gpurun_out/r02_n2_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02_n2_bench.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','bk1_ms','bk2_ms')})
print('e2e', d['e2e']['value'], d['e2e']['two_call_value'], d['e2e']['host_link_probe'])
for c in d['configs']: print({k:v for k,v in c.items() if k in ('name','bk1_states_per_s','bk2_states_per_s','states_per_s','error')})
"
