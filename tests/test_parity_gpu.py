"""GPU parity tests: the CUDA path, called through the C ABI (bbduk_b200_process / _process_device),
against the CPU oracle on the same seeded inputs. Bit-exact on every output array, the aggregate
counters, the per-scaffold hit counts and the table itself."""
import numpy as np
import pytest

from bbtools_b200 import make_cfg, synth
from bbtools_b200.fasta import pack

pytestmark = pytest.mark.gpu


def engines(adapters, ref=None, **kw):
    from bbtools_b200.bbduk import BBDukIndexGPU
    from oracle.oracle import Oracle
    cfg = make_cfg(**kw)
    _, b, off = adapters if ref is None else (None, *ref)
    o = Oracle(cfg)
    o.add_ref(b, off)
    n_o = o.finalize()
    g = BBDukIndexGPU(cfg)
    g.add_ref(b, off)
    n_g = g.finalize()
    assert n_g == n_o, f"stored k-mers differ: gpu {n_g} oracle {n_o}"
    return o, g


def assert_same(o, g, bases, offsets, paired, want_mask=False, threads=8):
    eo, so = o.process(bases, offsets, paired, threads=threads, want_mask=want_mask)
    eg, sg = g.process(bases, offsets, paired, want_mask=want_mask)
    for name, x in eo.fields().items():
        y = eg.fields()[name]
        if not np.array_equal(x, y):
            bad = np.nonzero(x != y)[0]
            i = int(bad[0])
            raise AssertionError(f"{name}: {len(bad)} mismatches, first at {i}: oracle {x[i]} gpu {y[i]}")
    assert so.as_dict() == sg.as_dict()
    return eo, so


def check_scaffold_counts(o, g):
    ro, bo = o.scaffold_counts()
    rg, bg = g.scaffold_counts()
    assert np.array_equal(ro, rg) and np.array_equal(bo, bg)


MODES = [
    dict(k=23, ktrim_right=1),                                              # cfg 1
    dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1),       # cfg 2 (headline)
    dict(k=23, mink=11, hdist=1, ktrim_left=1),
    dict(k=23, mink=11, hdist=1, ktrim_left=1, ktrim_right=1),              # tips
    dict(k=23, mink=8, hdist=1, hdist2=0, ktrim_right=1, ktrim_exclusive=1, trim_pad=2),
    dict(k=25, hdist=1, ktrim_left=1, ktrim_right=1, restrict_left=60, restrict_right=60),
    dict(k=27, hdist=1, ktrim_right=1, restrict_right=50),
    dict(k=21, hdist=0, ktrim_right=1, forbid_ns=1, mask_middle=0, min_len_fraction=0.5),
    dict(k=23, mink=11, hdist=1, ktrim_right=1, skip_r2=1, require_both_bad=1),
    dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_failures_to_1bp=1, trim_pairs_evenly=1),
    dict(k=31),                                                             # kfilter, cfg-3 shape
    dict(k=31, max_bad_kmers=2, mask_middle=0),
    dict(k=27, hdist=1, min_kmer_fraction=0.05),
    dict(k=25, min_covered_fraction=0.2),
    dict(k=25, find_best_match=1),
    dict(k=40),                                                             # k>31 -> countSetKmersBig
    dict(k=23, qhdist=1, ktrim_right=1),
    dict(k=20, speed=5),
    dict(k=20, speed=5, generation=1),                                      # bbduk.BBDukS speed= rule (hash bits 16-19)
    dict(k=23, speed=9, generation=1, ktrim_right=1),
    dict(k=20, qskip=3, ktrim_right=1),
    dict(k=19, rcomp=0, ktrim_right=1, mid_mask_len=3),
    dict(k=27, hdist=2, ktrim_right=1),                                     # cfg 4 table (8.4 M keys)
]


@pytest.mark.parametrize("kw", MODES, ids=lambda kw: ",".join(f"{a}={b}" for a, b in kw.items()))
def test_modes_paired_and_ragged(adapters, adapter_seqs, kw):
    o, g = engines(adapters, **kw)
    b, off = synth.paired_adapter_reads(6000, seed=3)
    assert_same(o, g, b, off, True)
    assert_same(o, g, b, off, False)
    rb, ro = synth.ragged_reads(3000, seed=4, adapter=adapter_seqs[0].encode())
    assert_same(o, g, rb, ro, False)
    rb, ro = synth.ragged_reads(3000, seed=5, adapter=adapter_seqs[2].encode(), max_len=120)
    assert_same(o, g, rb, ro, True)
    check_scaffold_counts(o, g)


@pytest.mark.parametrize("kw", [dict(k=23, mink=11, hdist=1, ktrim_right=1), dict(k=23, mink=11, hdist=1, ktrim_left=1),
                                dict(k=23, hdist=1), dict(k=25, hdist=2, ktrim_right=1), dict(k=21, hdist=1, rcomp=0, ktrim_right=1),
                                dict(k=23, hdist=1, forbid_ns=1, ktrim_right=1)],
                         ids=lambda kw: ",".join(f"{a}={b}" for a, b in kw.items()))
def test_undefined_bases_inside_adapters(adapter_seqs, adapters, kw):
    """Without forbidNs the forward k-mer reads an undefined base as A and the reverse k-mer as the complement of T
    (jgi/BBDuk.java:3882-3888), so a window with an N can hit through either reading. Every position of an adapter
    (and of its reverse complement) is replaced by N / another IUPAC code in turn, alone, as a pair 1-12 bases apart
    and together with one substitution, in front of and behind random flanks."""
    o, g = engines(adapters, **kw)
    rng = np.random.default_rng(11)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    for seq in (adapter_seqs[0], adapter_seqs[2], adapter_seqs[5]):
        fwd = seq.encode()[:40]
        for ad in (fwd, fwd.translate(comp)[::-1]):
            for i in range(len(ad)):
                for gap in (0, 1, 5, 12):
                    a = bytearray(ad)
                    a[i] = ord("N")
                    if gap and i + gap < len(a):
                        a[i + gap] = rng.choice(np.frombuffer(b"NRYn", np.uint8))
                    if gap == 5 and i >= 3:
                        a[i - 3] = ord("ACGT"[(b"ACGT".index(bytes([ad[i - 3]])) + 1) % 4])
                    left = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(0, 60))))
                    right = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(0, 30))))
                    reads.append(left + bytes(a) + right)
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    b = np.frombuffer(b"".join(reads), np.uint8).copy()
    eo, _ = assert_same(o, g, b, off, False)
    assert_same(o, g, b, off, True)
    if not kw.get("forbid_ns"):
        assert int((eo.fields()["id0"] > 0).sum()) > len(reads) // 4  # the cases do hit


KMASK = [
    dict(k=23, ktrim_n=1),
    dict(k=23, mink=11, hdist=1, ktrim_n=1),
    dict(k=23, mink=11, hdist=1, ktrim_n=1, kmask_fully_covered=1),
    dict(k=21, hdist=1, ktrim_n=1, trim_pad=3, restrict_left=70),
]


@pytest.mark.parametrize("kw", KMASK, ids=lambda kw: ",".join(f"{a}={b}" for a, b in kw.items()))
def test_kmask(adapters, adapter_seqs, kw):
    o, g = engines(adapters, **kw)
    b, off = synth.paired_adapter_reads(4000, seed=6)
    assert_same(o, g, b, off, True, want_mask=True)
    rb, ro = synth.ragged_reads(3000, seed=7, adapter=adapter_seqs[0].encode())
    assert_same(o, g, rb, ro, False, want_mask=True)
    check_scaffold_counts(o, g)


@pytest.mark.parametrize("kw", [dict(k=23, ksplit=1), dict(k=21, mink=9, hdist=1, ksplit=1)],
                         ids=lambda kw: ",".join(f"{a}={b}" for a, b in kw.items()))
def test_ksplit(adapters, adapter_seqs, kw):
    o, g = engines(adapters, **kw)
    rb, ro = synth.ragged_reads(4000, seed=8, adapter=adapter_seqs[0].encode(), min_len=0, max_len=300)
    assert_same(o, g, rb, ro, False)
    b, off = synth.single_adapter_reads(5000, adapter_seqs[0].encode()[:40], seed=2, min_off=60)
    assert_same(o, g, b, off, False)


@pytest.mark.parametrize("kw", [dict(k=23, ktrim_right=1), dict(k=23, mink=11, hdist=1, ktrim_right=1),
                                dict(k=21, hdist=1, edist=1, ktrim_right=1), dict(k=15, edist=2, hdist2=1, mink=9, ktrim_left=1),
                                dict(k=27, hdist=2, ktrim_right=1), dict(k=23, hdist=0, speed=7),
                                dict(k=19, hdist=3, mask_middle=0, ktrim_right=1)],
                         ids=lambda kw: ",".join(f"{a}={b}" for a, b in kw.items()))
def test_table_is_identical(adapters, kw):
    """every key and every id, including edit-distance neighbourhoods and min-id ties"""
    _, b, off = adapters
    if kw.get("hdist") == 3 or kw.get("edist") == 2:
        b, off = b[:off[12]], off[:13]  # keep the literal recursion of the oracle affordable
    o, g = engines(None, ref=(b, off), **kw)
    ko, vo = o.dump_table()
    kg, vg = g.dump_table()
    assert np.array_equal(ko, kg)
    assert np.array_equal(vo, vg)


def test_reference_edge_cases():
    """scaffolds shorter than k, N / IUPAC inside scaffolds, lowercase, empty scaffold, rskip"""
    seqs = [b"ACGT", b"", b"ACGTNACGTTGCATGGATCCAGTACGATTACAGGCATRACGATCAGCATCGACTAGCATCGACTAG",
            b"acgattagcgcgcgattttagagagctctcgagagcttcgagagctcttgaga", b"GATTACAGATTACAGATTACAGATTACAGATTACA" * 40]
    ref = pack(seqs)
    for kw in (dict(k=11, ktrim_right=1), dict(k=13, mink=6, hdist=1, ktrim_right=1), dict(k=11, min_skip=3, max_skip=3),
               dict(k=31, hdist=1)):
        o, g = engines(None, ref=ref, **kw)
        assert np.array_equal(o.dump_table()[0], g.dump_table()[0])
        assert np.array_equal(o.dump_table()[1], g.dump_table()[1])
        rb, ro = synth.ragged_reads(2000, seed=1, adapter=seqs[3].upper(), max_len=90)
        assert_same(o, g, rb, ro, False)


def test_empty_and_degenerate_batches(adapters):
    o, g = engines(adapters, k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
    for seqs, paired in (([], False), ([b""], False), ([b"", b""], True), ([b"A"], False),
                         ([b"N" * 200, b"ACGT" * 50], True), ([b"AGATCGGAAGAGC", b"AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"], True)):
        b, off = pack(seqs)
        assert_same(o, g, b, off, paired)
    # empty reference: nothing happens to any read
    from bbtools_b200.bbduk import BBDukIndexGPU
    from oracle.oracle import Oracle
    cfg = make_cfg(k=23, ktrim_right=1)
    o2, g2 = Oracle(cfg), BBDukIndexGPU(cfg)
    assert o2.finalize() == 0 and g2.finalize() == 0
    b, off = synth.paired_adapter_reads(100)
    assert_same(o2, g2, b, off, True)


def test_long_reads_and_chunking(adapters, adapter_seqs):
    """reads far beyond the staged fast-path length, and a batch that spans several device chunks"""
    o, g = engines(adapters, k=23, mink=11, hdist=1, ktrim_right=1)
    rb, ro = synth.ragged_reads(300, seed=12, adapter=adapter_seqs[0].encode(), min_len=1000, max_len=20000)
    assert_same(o, g, rb, ro, False)
    o, g = engines(adapters, k=27, ktrim_left=1)
    assert_same(o, g, rb, ro, True)


def test_device_resident_path_and_device_synth(adapters):
    """bbduk_b200_process_device on HBM-resident buffers + the device generator equals numpy's"""
    import ctypes as C

    import torch

    from bbtools_b200 import _lib
    o, g = engines(adapters, k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
    n_pairs, L = 20000, 150
    d_bases = torch.empty(2 * n_pairs * L, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(2 * n_pairs + 1, dtype=torch.int32, device="cuda")
    lib = _lib.load()
    rc = lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), n_pairs, 1000, L, C.c_uint64(1), 50, 5, None)
    assert rc == 0
    torch.cuda.synchronize()
    hb, ho = synth.paired_adapter_reads(n_pairs, first_pair=1000, read_len=L, seed=1)
    assert np.array_equal(d_bases.cpu().numpy(), hb)
    assert np.array_equal(d_off.cpu().numpy().astype(np.int64), ho)
    n = 2 * n_pairs
    outs = {k: torch.empty(n, dtype=torch.int32, device="cuda") for k in ("id0", "hi", "lo", "count", "id0b")}
    outs["flags"] = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    g.process_device(d_bases, d_off, n, True, outs, d_stats=d_stats)
    torch.cuda.synchronize()
    eo, so = o.process(hb, ho, True, threads=8)
    for name in ("id0", "id0b", "lo", "hi", "count", "flags"):
        assert np.array_equal(outs[name].cpu().numpy(), eo.fields()[name]), name
    assert d_stats.cpu().tolist() == list(so.as_dict().values())


def test_concurrent_callers_share_one_table(adapters):
    """the reference calls its index from many ProcessThreads at once (bbduk/BBDukS.java:317-319)"""
    import threading
    o, g = engines(adapters, k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
    batches = [synth.paired_adapter_reads(3000, first_pair=3000 * i, seed=21) for i in range(6)]
    want = [o.process(b, off, True)[0] for b, off in batches]
    got = [None] * len(batches)

    def work(i):
        got[i] = g.process(batches[i][0], batches[i][1], True)[0]
    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(batches))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for w, x in zip(want, got):
        for name, a in w.fields().items():
            assert np.array_equal(a, x.fields()[name]), name


@pytest.mark.parametrize("kw", [dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1), dict(k=23, mink=11, hdist=1, ktrim_left=1),
                                dict(k=31), dict(k=25, find_best_match=1), dict(k=23, mink=11, hdist=1, ktrim_left=1, ktrim_right=1)],
                         ids=lambda kw: ",".join(f"{a}={b}" for a, b in kw.items()))
def test_packed_host_input(adapters, adapter_seqs, kw):
    """bbduk_b200_process_packed: the caller's own 2-bit stream + defined bits (as bbduk_b200_pack_bases writes them) give the
    oracle's results, for the tuned kernels (which read the stream as it is) and for the other modes (spelled out again on the
    device), on multi-chunk batches whose chunks start in the middle of a 16-base group."""
    o, g = engines(adapters, **kw)
    for (b, off), paired in ((synth.paired_adapter_reads(6000, seed=13), True),
                             (synth.ragged_reads(5000, seed=14, adapter=adapter_seqs[1].encode()), False)):
        eo, so = o.process(b, off, paired, threads=8)
        F, D = g.pack(b)
        eg, sg = g.process_packed(F, D, off, paired)
        for name, x in eo.fields().items():
            assert np.array_equal(x, eg.fields()[name]), name
        assert so.as_dict() == sg.as_dict()
    # 2.2 M ragged reads: several chunks (1 Mi reads each), chunk starts not aligned to a group
    rb, ro = synth.ragged_reads(40000, seed=15, adapter=adapter_seqs[0].encode(), max_len=90)
    reps = 56
    big = np.tile(rb, reps)
    boff = np.concatenate([[0], (ro[1:][None, :] + (np.arange(reps) * ro[-1])[:, None]).ravel()]).astype(np.int64)
    F, D = g.pack(big)
    eg, sg = g.process_packed(F, D, boff, False)
    ea, sa = g.process(big, boff, False)
    for name, x in ea.fields().items():
        assert np.array_equal(x, eg.fields()[name]), name
    assert sa.as_dict() == sg.as_dict()
    eo, _ = o.process(rb, ro, False, threads=8)
    assert np.array_equal(eg.fields()["hi"][-len(ro) + 1:], eo.fields()["hi"])


def test_packed_host_input_refuses_kmask(adapters):
    _, g = engines(adapters, k=23, ktrim_n=1)
    b, off = synth.paired_adapter_reads(100, seed=2)
    F, D = g.pack(b)
    with pytest.raises(RuntimeError, match="kmask"):
        g.process_packed(F, D, off, True)
