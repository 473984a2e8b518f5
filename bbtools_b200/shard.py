"""Read sharding across the GPUs of one box: contiguous 1/N slices of a batch, pairs never
split, results returned in input order. No data-path collective: only the (optional) gather of results to
rank 0 uses torch.distributed, and it works the same over gloo (CPU tests) and NCCL."""
import numpy as np


def shard_units(n_units: int, world: int, rank: int):
    """[start, stop) of rank's contiguous slice of n_units (the reference's analogue is one ListNum per
    ProcessThread, bbduk/BBDukS.java:317-319)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n_units * rank // world, n_units * (rank + 1) // world


def shard_reads(offsets: np.ndarray, paired: bool, world: int, rank: int):
    """-> (first_read, last_read_exclusive, local_offsets rebased to 0)"""
    n_reads = len(offsets) - 1
    per = 2 if paired else 1
    if n_reads % per:
        raise ValueError("paired input needs an even number of reads")
    u0, u1 = shard_units(n_reads // per, world, rank)
    r0, r1 = u0 * per, u1 * per
    return r0, r1, (offsets[r0:r1 + 1] - offsets[r0]).astype(np.int64)


def gather_in_order(local: dict, group=None):
    """all ranks contribute {name: ndarray}; every rank gets the rank-order concatenation back."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    return {k: np.concatenate([p[k] for p in parts]) for k in local}


def sum_stats(stats: dict, group=None):
    """additive counters summed over ranks, as the reference sums its per-thread counters
    (jgi/BBDuk.java:2085-2131)."""
    import torch
    import torch.distributed as dist
    keys = sorted(stats)
    t = torch.tensor([int(stats[k]) for k in keys], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    return dict(zip(keys, t.cpu().tolist()))
