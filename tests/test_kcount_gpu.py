"""GPU parity of the KmerCountExact counting path (include/kcount_b200.h) against the oracle, through the C ABI.
Bit-exact: the multiset of (key,count), Unique Kmers, kmersIn and the count histogram."""
import numpy as np
import pytest
import torch

from bbtools_b200 import _lib, synth
from bbtools_b200.kcount import KmerTableSetGPU
from oracle.kcount import KCountOracle

pytestmark = pytest.mark.gpu


def same_tables(gpu, ora, histmax=1000):
    gk, gc = gpu.dump()
    order = np.argsort(gk)
    ok_, oc = ora.dump()
    assert np.array_equal(gk[order], ok_), "key sets differ"
    assert np.array_equal(gc[order], oc), "counts differ"
    assert gpu.stats() == ora.stats()
    assert np.array_equal(gpu.khist(histmax), ora.khist(histmax))


@pytest.mark.parametrize("k,rcomp", [(31, True), (21, True), (16, True), (5, True), (1, True), (31, False), (17, False)])
def test_ragged_reads(adapters, k, rcomp):
    _, ab, _ = adapters
    bases, offsets = synth.ragged_reads(4000, seed=100 + k, max_len=300, adapter=bytes(ab[:80]))
    g, o = KmerTableSetGPU(k, rcomp), KCountOracle(k, rcomp)
    g.add_reads(bases, offsets)
    o.add_reads(bases, offsets)
    same_tables(g, o)


def test_genome_reads_grow_from_a_tiny_table():
    bases, offsets = synth.genome_reads(60000, 200000, seed=11, sub_per_10k=10)
    g, o = KmerTableSetGPU(31), KCountOracle(31)
    n0 = g.table_info()["n_slots"]
    for a in range(0, 60000, 20000):  # three calls accumulate
        sl = slice(offsets[a], offsets[a + 20000])
        g.add_reads(bases[sl], offsets[a:a + 20001] - offsets[a])
        o.add_reads(bases[sl], offsets[a:a + 20001] - offsets[a])
    assert g.table_info()["n_slots"] > n0  # it had to resize
    same_tables(g, o, histmax=100000)
    assert g.stats()["kmers_in"] == 60000 * 120


def test_long_contig_short_and_empty_reads():
    contig = synth.genome_bases(1_000_003, seed=5)
    contig[500000:500040] = ord("N")
    contig[777] = ord("n")
    parts = [contig, np.zeros(0, np.uint8), np.frombuffer(b"ACGTACGTAC", np.uint8), np.zeros(0, np.uint8),
             np.frombuffer(b"ACGTTGCAACGTTGCAACGTTGCAACGTTGCAACG", np.uint8), contig[:4097], contig[10:31], contig[10:40]]
    bases = np.concatenate(parts)
    offsets = np.zeros(len(parts) + 1, np.int64)
    np.cumsum([len(p) for p in parts], out=offsets[1:])
    for k in (31, 11):
        g, o = KmerTableSetGPU(k), KCountOracle(k)
        g.add_reads(bases, offsets)
        o.add_reads(bases, offsets)
        same_tables(g, o)


def test_empty_batches():
    g = KmerTableSetGPU(31)
    g.add_reads(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    g.add_reads(np.zeros(0, np.uint8), np.zeros(5, np.int64))
    st = g.stats()
    assert st["unique_kmers"] == 0 and st["kmers_in"] == 0 and st["reads_in"] == 4
    assert g.khist(10).sum() == 0


def test_device_generator_and_device_entry_point():
    lib = _lib.load()
    n, L, G = 50000, 150, 300000
    d_bases = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    assert lib.kcount_b200_synth_reads(d_bases.data_ptr(), d_off.data_ptr(), n, 1000, L, G, 11, 10, None) == 0
    torch.cuda.synchronize()
    hb, ho = synth.genome_reads(n, G, first_read=1000, read_len=L, seed=11, sub_per_10k=10)
    assert np.array_equal(d_bases.cpu().numpy(), hb)
    assert np.array_equal(d_off.cpu().numpy().astype(np.int64), ho)
    g, o = KmerTableSetGPU(31, initial_keys=1 << 20), KCountOracle(31)
    g.add_reads_device(d_bases, d_off, n, n * L)
    o.add_reads(hb, ho)
    same_tables(g, o, histmax=100000)


def test_saturation_merge_and_partitioned_export():
    g = KmerTableSetGPU(31)
    keys = torch.tensor([7, 9, 11], dtype=torch.int64, device="cuda")
    g.merge(keys, torch.tensor([0x7FFFFFF0, 5, 1], dtype=torch.int32, device="cuda"))
    g.merge(keys, torch.tensor([100, 6, 0x7FFFFFFF], dtype=torch.int32, device="cuda"))
    k_, c_ = g.dump()
    order = np.argsort(k_)
    assert list(k_[order]) == [7, 9, 11] and list(c_[order]) == [0x7FFFFFFF, 11, 0x7FFFFFFF]
    # export by owner, merge the parts into a fresh table: same content
    bases, offsets = synth.genome_reads(20000, 100000, seed=3, sub_per_10k=20)
    a, o = KmerTableSetGPU(25), KCountOracle(25)
    a.add_reads(bases, offsets)
    o.add_reads(bases, offsets)
    ek, ec, sizes = a.export_partitioned(3)
    assert sum(sizes) == a.stats()["unique_kmers"] and min(sizes) > 0
    owner = (synth.mix64(ek.cpu().numpy().view(np.uint64)) >> np.uint64(32)) % np.uint64(3)
    assert np.array_equal(owner, np.repeat(np.arange(3, dtype=np.uint64), sizes))
    b = a.new_like()
    s0 = 0
    for s in sizes:
        b.merge(ek[s0:s0 + s].contiguous(), ec[s0:s0 + s].contiguous())
        s0 += s
    bk, bc = b.dump()
    order = np.argsort(bk)
    ok_, oc = o.dump()
    assert np.array_equal(bk[order], ok_) and np.array_equal(bc[order], oc)
