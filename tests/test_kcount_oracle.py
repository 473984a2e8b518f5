"""CPU tests of the KmerCountExact oracle (oracle/kcount_oracle.c): the serial port of
kmer/KmerTableSet.java:652-716 against an independent numpy restatement (explicit enumeration of every
window of k defined bases, canonical max, np.unique), plus the world_size-2 gloo test of the exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bbtools_b200 import synth
from oracle.kcount import KCountOracle

CODE = np.full(256, -1, np.int64)
for i, ch in enumerate("ACGT"):
    CODE[ord(ch)] = CODE[ord(ch.lower())] = i
CODE[ord("U")] = CODE[ord("u")] = 3


def numpy_counts(bases, offsets, k, rcomp=True):
    """independent restatement: per read, all windows of k defined bases -> canonical key -> counts"""
    keys = []
    n_kmers = 0
    for r in range(len(offsets) - 1):
        c = CODE[bases[offsets[r]:offsets[r + 1]]]
        L = len(c)
        if L < k:
            continue
        ok = np.convolve((c >= 0).astype(np.int64), np.ones(k, np.int64), "valid") == k  # window start s
        idx = np.nonzero(ok)[0]
        if len(idx) == 0:
            continue
        win = c[idx[:, None] + np.arange(k)[None, :]].astype(object)
        fwd = np.zeros(len(idx), object)
        rev = np.zeros(len(idx), object)
        for j in range(k):
            fwd = fwd * 4 + win[:, j]
            rev = rev * 4 + (3 - win[:, k - 1 - j])
        key = np.where(rev > fwd, rev, fwd) if rcomp else fwd
        keys.extend(int(x) for x in key)
        n_kmers += len(idx)
    u, cnt = np.unique(np.array(keys, np.uint64), return_counts=True)
    return u, cnt.astype(np.int64), n_kmers


@pytest.mark.parametrize("k,rcomp", [(31, True), (21, True), (5, True), (1, True), (31, False), (16, True)])
def test_oracle_matches_numpy(adapters, k, rcomp):
    _, ab, _ = adapters
    bases, offsets = synth.ragged_reads(300, seed=5 + k, max_len=120, adapter=bytes(ab[:60]))
    o = KCountOracle(k, rcomp)
    o.add_reads(bases, offsets)
    u, cnt, n_kmers = numpy_counts(bases, offsets, k, rcomp)
    keys, counts = o.dump()
    assert np.array_equal(keys, u) and np.array_equal(counts, cnt)
    st = o.stats()
    assert st["kmers_in"] == n_kmers and st["unique_kmers"] == len(u)
    assert st["reads_in"] == 300 and st["bases_in"] == offsets[-1]
    hist = o.khist(8)
    want = np.bincount(np.minimum(cnt, 8), minlength=9)
    assert np.array_equal(hist, want)


def test_genome_reads_counts_are_sane():
    bases, offsets = synth.genome_reads(2000, 5000, seed=11, sub_per_10k=10)
    o = KCountOracle(31, True)
    o.add_reads(bases, offsets)
    st = o.stats()
    assert st["kmers_in"] == 2000 * 120
    # 60x coverage of a 5 kbp genome: the genomic k-mers dominate, error k-mers are singletons
    hist = o.khist(1000)
    assert 4000 < st["unique_kmers"] < 20000 and hist[1] > 0 and hist[30:].sum() > 3000


def test_saturation_and_merge():
    o = KCountOracle(3, True)
    o.merge_arrays(np.array([7, 9], np.uint64), np.array([0x7FFFFFF0, 5], np.int32))
    o.merge_arrays(np.array([7, 9], np.uint64), np.array([100, 6], np.int32))
    keys, counts = o.dump()
    assert list(keys) == [7, 9] and list(counts) == [0x7FFFFFFF, 11]


# ---- the exchange (bbtools_b200.kcount.exchange_counts) with CPU stand-ins over gloo ------------------
class CpuTable:
    """same three methods as KmerTableSetGPU, backed by the oracle and CPU tensors"""

    def __init__(self, k=31, rcomp=True):
        self.o = KCountOracle(k, rcomp)

    def export_partitioned(self, n_parts):
        keys, counts = self.o.dump()
        owner = (synth.mix64(keys) >> np.uint64(32)) % np.uint64(n_parts)
        order = np.argsort(owner, kind="stable")
        sizes = np.bincount(owner.astype(np.int64), minlength=n_parts)
        return (torch.from_numpy(keys[order].view(np.int64).copy()), torch.from_numpy(counts[order].copy()),
                [int(x) for x in sizes])

    def new_like(self, initial_keys=0):
        return CpuTable(self.o.k, self.o.rcomp)

    def merge(self, keys, counts):
        self.o.merge_arrays(keys.numpy().view(np.uint64), counts.numpy())

    def khist(self, histmax):
        return self.o.khist(histmax)

    def stats(self):
        return self.o.stats()


def _worker(rank, world, port, n_reads, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bbtools_b200.kcount import exchange_counts, global_summary
    from bbtools_b200.shard import shard_reads
    bases, offsets = synth.genome_reads(n_reads, 3000, seed=11, sub_per_10k=30)
    r0, r1, loff = shard_reads(offsets, False, world, rank)
    t = CpuTable(21)
    t.o.add_reads(bases[offsets[r0]:offsets[r1]], loff)
    owner = exchange_counts(t)
    keys, _ = owner.o.dump()
    own = (synth.mix64(keys) >> np.uint64(32)) % np.uint64(world)
    assert np.all(own == rank)  # every key landed on its owner
    uniq, hist = global_summary(owner, 500)
    if rank == 0:
        q.put((uniq, hist))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_two_ranks_equal_one():
    n_reads = 1201
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_reads, q)) for r in range(2)]
    [p.start() for p in procs]
    uniq, hist = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    bases, offsets = synth.genome_reads(n_reads, 3000, seed=11, sub_per_10k=30)
    o = KCountOracle(21)
    o.add_reads(bases, offsets)
    assert uniq == o.stats()["unique_kmers"]
    assert np.array_equal(hist, o.khist(500))
