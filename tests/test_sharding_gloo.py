"""world_size-2 gloo test of the N>1 path's host logic: each rank takes its contiguous shard of the batch,
processes it independently (here with the oracle port standing in for the kernel) and a host-side gather
reproduces the single-rank answer exactly.  No collective is needed on the data path."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_states, out_path):
    sys.path.insert(0, ROOT)
    from kinetix_b200.sharding import shard_of
    from oracle.port import Port, R, synthetic_states
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    p = Port('LiDryer', transport=False)
    full = synthetic_states(p.N, n_states, seed=42)          # same seeded batch on every rank
    b, e = shard_of(n_states, rank, world)
    mine = p.production_rates(full[:, b:e], 101325.0 / R, 101325.0)
    # host-side gather for verification (variable shard sizes -> gather objects)
    gathered = [None] * world
    dist.all_gather_object(gathered, (b, e, mine))
    dist.barrier()
    if rank == 0:
        out = np.empty_like(full)
        for gb, ge, part in gathered:
            out[:, gb:ge] = part
        np.save(out_path, out)
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_single_rank(tmp_path):
    from oracle.port import Port, R, synthetic_states
    n_states = 1001                                          # odd: shards of 501 and 500
    out_path = str(tmp_path / 'gathered.npy')
    mp.spawn(_worker, args=(2, _free_port(), n_states, out_path), nprocs=2, join=True)
    p = Port('LiDryer', transport=False)
    full = synthetic_states(p.N, n_states, seed=42)
    ref = p.production_rates(full, 101325.0 / R, 101325.0)
    assert np.array_equal(np.load(out_path), ref)
