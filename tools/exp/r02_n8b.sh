#!/bin/bash
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n tools/e2e_phase_probe.py 2>/dev/null | tail -1 | tee -a gpurun_out/r02_e2e_phase_probe.txt
done
timeout 200 python tools/e2e_phase_probe.py 2>/dev/null | tail -1 | tee -a gpurun_out/r02_e2e_phase_probe.txt
