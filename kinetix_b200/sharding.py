"""State sharding across GPUs / ranks.

States are independent, so the batch is cut into contiguous ranges, one per rank, with no collective on
the data path.  The reference does `n_states /= size` and silently drops the remainder
(reference benchmark/src/bk.cpp:443); here the remainder is spread over the first ranks.
"""


def shard_bounds(n_states, world_size):
    """[(begin, end)] per rank: contiguous, disjoint, covering [0, n_states), sizes differ by at most 1."""
    if world_size <= 0:
        raise ValueError('world_size must be positive')
    base, extra = divmod(int(n_states), int(world_size))
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def shard_of(n_states, rank, world_size):
    return shard_bounds(n_states, world_size)[rank]
