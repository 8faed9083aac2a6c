"""Partition of the (frame pair, timestep) work items over ranks.

Every output pixel of every (pair, timestep) depends only on that pair's two frames and its
flows, so the path shards with no data-path collective.  All N timesteps of a pair stay on one
rank whenever there are at least as many pairs as ranks (stage 1 then runs once per pair and
both kernels get their timestep batching); otherwise the timesteps of each pair are split.
"""


def _block(n_items, n_parts, part):
    """[start, end) of block `part` when n_items are split into n_parts nearly equal blocks."""
    base, rem = divmod(n_items, n_parts)
    start = part * base + min(part, rem)
    return start, start + base + (1 if part < rem else 0)


def shard_work(n_pairs, n_timesteps, rank, world_size):
    """Work of `rank`: a list of (pair, t_start, t_end) with t_end exclusive; deterministic and
    exhaustive over ranks, no overlaps."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank %r / world_size %r" % (rank, world_size))
    if n_pairs >= world_size:
        p0, p1 = _block(n_pairs, world_size, rank)
        return [(p, 0, n_timesteps) for p in range(p0, p1)]
    # fewer pairs than ranks: flatten (pair, timestep) and block-distribute
    lo, hi = _block(n_pairs * n_timesteps, world_size, rank)
    work = []
    for p in range(n_pairs):
        a, b = max(lo, p * n_timesteps), min(hi, (p + 1) * n_timesteps)
        if a < b:
            work.append((p, a - p * n_timesteps, b - p * n_timesteps))
    return work


def frames_of(work):
    return sum(t1 - t0 for _, t0, t1 in work)
