#!/usr/bin/env python3
"""Host-link probe for the end-to-end path on a multi-GPU host (run under torchrun, one rank per GPU): the byte
streams of one BK1+BK2 host call (432 B up, 872 B down per GRI-3.0 state) moved (a) the way the library pipelines them
-- chunks on 4 streams, uploads and downloads in flight together -- and (b) in phases -- everything up, then everything
down -- with no kernels in between.  Prints the aggregate states/s each schedule would allow."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
S, N = 1 << 22, 53
up, down = (N + 1) * 8, (2 * N + 3) * 8
h_in = torch.empty(S * up, dtype=torch.uint8).pin_memory()
h_out = torch.empty(S * down, dtype=torch.uint8).pin_memory()
h_in.zero_(); h_out.zero_()
d_in = torch.empty(S * up, dtype=torch.uint8, device='cuda')
d_out = torch.empty(S * down, dtype=torch.uint8, device='cuda')
streams = [torch.cuda.Stream() for _ in range(4)]
CH = 1 << 17


def overlapped():
    for i, s0 in enumerate(range(0, S, CH)):
        with torch.cuda.stream(streams[i % 4]):
            d_in[s0 * up:(s0 + CH) * up].copy_(h_in[s0 * up:(s0 + CH) * up], non_blocking=True)
            h_out[s0 * down:(s0 + CH) * down].copy_(d_out[s0 * down:(s0 + CH) * down], non_blocking=True)
    torch.cuda.synchronize()


def phased(parts=1):
    P = S // parts
    for q in range(parts):
        d_in[q * P * up:(q + 1) * P * up].copy_(h_in[q * P * up:(q + 1) * P * up], non_blocking=True)
        h_out[q * P * down:(q + 1) * P * down].copy_(d_out[q * P * down:(q + 1) * P * down], non_blocking=True)
    torch.cuda.synchronize()


def timeit(fn, reps=3):
    fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t = torch.tensor([time.perf_counter() - t0], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return S * world * reps / float(t[0])


res = {'overlapped_4_streams_128Ki_chunks': timeit(overlapped), 'phased_whole_batch': timeit(phased),
       'phased_4_parts': timeit(lambda: phased(4)), 'phased_16_parts': timeit(lambda: phased(16))}
if rank == 0:
    print({'n_gpus': world, 'states_per_rank': S, **{k: f'{v:.3e} states/s = {v * (up + down) / 1e9:.1f} GB/s both ways' for k, v in res.items()}})
if world > 1:
    dist.destroy_process_group()
