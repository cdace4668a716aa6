#!/bin/bash
# 8-GPU call: bench at N=8 and N=4 (one process per GPU, weak scaling), reference arm at N=8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02_n8_gpus.txt
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_n${n}_bench.json 2> gpurun_out/r02_n${n}_bench.err; echo "bench N=$n rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02_n${n}_bench.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','bk1_ms','bk2_ms')})
print('e2e', d['e2e']['value'], d['e2e']['two_call_value'], d['e2e']['host_link_probe'])
for c in d['configs']: print({k:v for k,v in c.items() if k in ('name','bk1_states_per_s','bk2_states_per_s','states_per_s','error')})
"
done
