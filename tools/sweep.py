#!/usr/bin/env python3
"""Precision / size sweep of BK1 and BK2 on ONE GPU (BASELINE.json config 5, and the per-GPU points of configs
3 and 4).  Under torchrun each rank runs the same sizes on its own shard (weak scaling) and rank 0 reports
the aggregate.  Output: JSON lines (one per mechanism x mode x size) on stdout / --out.

  python tools/sweep.py --mech gri30 --modes f64,fpmix,f32 --min 1024 --max 134217728 --out profiles/sweep_r01.jsonl
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinetix_b200.host as kinetix  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--mech', default='gri30')
ap.add_argument('--modes', default='f64,fpmix,f32')
ap.add_argument('--min', type=int, default=1024)
ap.add_argument('--max', type=int, default=1 << 27)
ap.add_argument('--factor', type=int, default=4)
ap.add_argument('--out', default=None)
ap.add_argument('--kernels', default='bk1,bk2')
a = ap.parse_args()

rank = int(os.environ.get('RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

lines = []
for mode in a.modes.split(','):
    sp = mode != 'f64'
    dtype = 1 if mode == 'f32' else 0
    tdt = torch.float32 if mode == 'f32' else torch.float64
    kinetix.init(os.path.join(ROOT, 'kinetix_b200', 'mechanisms', a.mech + '.yaml'), device_id=local,
                 single_precision=sp)
    N = kinetix.nSpecies()
    kinetix.build(101325.0, 1.0, [1.0 / N] * N, True)
    S = a.min
    while S <= a.max:
        need = 2 * (N + 1) * S * (4 if mode == 'f32' else 8)
        if need > 0.9 * torch.cuda.mem_get_info()[1]:
            break
        gen = torch.Generator(device='cuda')
        gen.manual_seed(1234 + rank)
        st = torch.empty((N + 1, S), dtype=tdt, device='cuda')
        st[0].uniform_(300.0, 2500.0, generator=gen)
        st[1:].uniform_(0.0, 1.0, generator=gen)
        st[1:] /= st[1:].sum(dim=0, keepdim=True)
        out = torch.empty_like(st)
        rec = dict(mechanism=a.mech, mode=mode, n_states_per_gpu=S, n_gpus=world)

        def timeit(fn):
            fn()
            fn()
            torch.cuda.synchronize()
            reps = max(3, min(200, int(2e8 / max(S * N, 1))))
            if dist is not None:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device='cuda', dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        if 'bk1' in a.kernels:
            ms = timeit(lambda: kinetix.productionRates(S, S, S, 1.0, st, out, dtype=dtype))
            rec.update(bk1_ms=ms, bk1_states_per_s=S * world / ms * 1e3)
        if 'bk2' in a.kernels:
            visc, cond, rhoD = out[0], out[1, :S].clone(), out[1:]
            ms = timeit(lambda: kinetix.mixtureAvgTransportProps(S, S, S, 1.0, st, visc, cond, rhoD[:N], dtype=dtype))
            rec.update(bk2_ms=ms, bk2_states_per_s=S * world / ms * 1e3)
        del st, out
        torch.cuda.empty_cache()
        if rank == 0:
            print(json.dumps(rec), flush=True)
            lines.append(rec)
        S *= a.factor
    kinetix.finalize()
if rank == 0 and a.out:
    with open(a.out, 'a') as fh:
        for rec in lines:
            fh.write(json.dumps(rec) + '\n')
if dist is not None:
    dist.destroy_process_group()
