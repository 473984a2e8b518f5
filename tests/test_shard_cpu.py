"""world_size-2 gloo test of the N>1 host logic: contiguous read shards, pairs never split, results
gathered in input order, counters summed -- checked against the single-process oracle run."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from bbtools_b200 import make_cfg, synth
from bbtools_b200.shard import gather_in_order, shard_reads, shard_units, sum_stats

CFG = dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)


def test_shard_units_cover_everything_once():
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_units(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bbtools_b200.fasta import read_fasta
    from oracle.oracle import Oracle  # the checker stands in for the per-GPU engine in this CPU test
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _, rb, roff = read_fasta(os.path.join(root, "tests", "golden", "adapters.fa"))
    bases, offsets = synth.paired_adapter_reads(n_pairs, seed=13)
    r0, r1, loff = shard_reads(offsets, True, world, rank)
    assert r0 % 2 == 0 and r1 % 2 == 0
    o = Oracle(make_cfg(**CFG))
    o.add_ref(rb, roff)
    o.finalize()
    out, st = o.process(bases[offsets[r0]:offsets[r1]], loff, True)
    merged = gather_in_order({k: v for k, v in out.fields().items()})
    total = sum_stats(st.as_dict())
    if rank == 0:
        q.put((merged, total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(adapters):
    from oracle.oracle import Oracle
    n_pairs = 1501  # odd on purpose: the shards are unequal
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    [p.start() for p in procs]
    merged, total = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    _, rb, roff = adapters
    o = Oracle(make_cfg(**CFG))
    o.add_ref(rb, roff)
    o.finalize()
    bases, offsets = synth.paired_adapter_reads(n_pairs, seed=13)
    want, st = o.process(bases, offsets, True)
    for k, v in want.fields().items():
        assert np.array_equal(v, merged[k]), k
    assert total == st.as_dict()
