"""GPU tests of table replication: a table that was NOT built on a handle but copied into it --
through table_describe -> table_alloc -> blob copies -> table_commit (what bbduk.BBDukIndexGPU.broadcast_table does with
NCCL under torchrun), and through the in-library bbduk_b200_replicate -- must answer every read like the oracle and hold
the same keys and ids as the source. Also the in-library read sharding (bbduk_b200_process_sharded)."""
import ctypes as C

import numpy as np
import pytest

from bbtools_b200 import make_cfg, synth

pytestmark = pytest.mark.gpu

CFG2 = dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)


def build(adapters, **kw):
    from bbtools_b200.bbduk import BBDukIndexGPU
    from oracle.oracle import Oracle
    cfg = make_cfg(**kw)
    _, b, off = adapters
    o = Oracle(cfg)
    o.add_ref(b, off)
    n_o = o.finalize()
    g = BBDukIndexGPU(cfg)
    g.add_ref(b, off)
    assert g.finalize() == n_o
    return cfg, o, g


def same_as_oracle(o, g, bases, offsets, paired):
    eo, so = o.process(bases, offsets, paired, threads=8)
    eg, sg = g.process(bases, offsets, paired)
    for name, x in eo.fields().items():
        y = eg.fields()[name]
        assert np.array_equal(x, y), f"{name}: {int((x != y).sum())} mismatches"
    assert so.as_dict() == sg.as_dict()


def test_table_alloc_commit_by_hand(adapters):
    """the three-call replication path of include/bbduk_b200.h with the blobs copied by the caller"""
    import torch
    from bbtools_b200._abi import BBDukTableDesc
    from bbtools_b200.bbduk import BBDukIndexGPU
    cfg, o, a = build(adapters, **CFG2)
    da = a.table_describe()
    b = BBDukIndexGPU(cfg)
    db = BBDukTableDesc()
    db.n_slots, db.n_filter_words, db.stored_kmers, db.n_scaffolds = da.n_slots, da.n_filter_words, da.stored_kmers, da.n_scaffolds
    for i in range(8):
        db.scalars[i] = da.scalars[i]
    b._check(b.lib.bbduk_b200_table_alloc(b.h, C.byref(db)), "table_alloc")
    assert db.d_keys and db.d_keys != da.d_keys
    for src, dst, nbytes in ((da.d_keys, db.d_keys, da.n_slots * 8), (da.d_vals, db.d_vals, da.n_slots * 4),
                             (da.d_filter, db.d_filter, da.n_filter_words * 4)):
        b._view(dst, nbytes).copy_(a._view(src, nbytes))
    torch.cuda.synchronize()
    b._check(b.lib.bbduk_b200_table_commit(b.h), "table_commit")
    b.stored_kmers, b.n_scaffolds = a.stored_kmers, a.n_scaffolds
    ka, va = a.dump_table()
    a.close()  # the copy must not depend on the source's memory
    kb, vb = b.dump_table()
    assert np.array_equal(ka, kb) and np.array_equal(va, vb) and len(kb) == 217135
    bases, offsets = synth.paired_adapter_reads(20000, seed=11)
    same_as_oracle(o, b, bases, offsets, True)
    ro, bo = o.scaffold_counts()
    rg, bg = b.scaffold_counts()
    assert np.array_equal(ro, rg) and np.array_equal(bo, bg)


@pytest.mark.parametrize("kw", [CFG2, dict(k=31), dict(k=27, hdist=2, ktrim_right=1)],
                         ids=["cfg2", "k31-kfilter", "k27-hdist2"])
def test_replicate_same_device(adapters, kw):
    """bbduk_b200_replicate onto the source's own GPU (peer-copy transport): two replicas, source destroyed first"""
    cfg, o, a = build(adapters, **kw)
    ka, va = a.dump_table()
    reps = a.replicate([0, 0])
    assert [r.transport for r in reps] == ["peer", "peer"] and a.transport == "built"
    a.close()
    bases, offsets = synth.paired_adapter_reads(8000, seed=12)
    for r in reps:
        kb, vb = r.dump_table()
        assert np.array_equal(ka, kb) and np.array_equal(va, vb)
    same_as_oracle(o, reps[0], bases, offsets, True)
    rb, ro_ = synth.ragged_reads(3000, seed=13)
    same_as_oracle(o, reps[1], rb, ro_, False)


def test_process_sharded_one_gpu(adapters):
    """the in-library sharding with three handles on one GPU: results in input order, counters and scaffold counts summed"""
    from bbtools_b200.bbduk import BBDukIndexGPU
    cfg, o, a = build(adapters, **CFG2)
    reps = [a] + a.replicate([0, 0])
    bases, offsets = synth.paired_adapter_reads(30001, seed=14)  # an odd pair count: uneven slices
    eo, so = o.process(bases, offsets, True, threads=8)
    eg, sg = BBDukIndexGPU.process_sharded(reps, bases, offsets, True)
    for name, x in eo.fields().items():
        assert np.array_equal(x, eg.fields()[name]), name
    assert so.as_dict() == sg.as_dict()
    ro, bo = o.scaffold_counts()
    rg, bg = BBDukIndexGPU.scaffold_counts_sum(reps)
    assert np.array_equal(ro, rg) and np.array_equal(bo, bg)
    # kmask through the sharded call: the mask words of all slices land in the caller's one buffer
    cfgm, om, am = build(adapters, k=23, mink=11, hdist=1, ktrim_n=1)
    repm = [am] + am.replicate([0])
    rb, ro_ = synth.ragged_reads(4001, seed=15)
    em, _ = om.process(rb, ro_, False, threads=8, want_mask=True)
    gm, _ = BBDukIndexGPU.process_sharded(repm, rb, ro_, False, want_mask=True)
    for name, x in em.fields().items():
        assert np.array_equal(x, gm.fields()[name]), name


def test_replicate_across_gpus(adapters):
    """needs >= 2 visible GPUs (gpurun --gpus 2): NCCL broadcast inside the library, then the batch sharded over both"""
    import torch
    from bbtools_b200.bbduk import BBDukIndexGPU
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one visible GPU")
    cfg, o, a = build(adapters, **CFG2)
    ka, va = a.dump_table()
    others = a.replicate(list(range(1, n)))
    assert all(r.transport == "nccl" for r in others), [r.transport for r in others]
    for i, r in enumerate(others):
        with torch.cuda.device(i + 1):
            kb, vb = r.dump_table()
        assert np.array_equal(ka, kb) and np.array_equal(va, vb)
    bases, offsets = synth.paired_adapter_reads(40000, seed=16)
    eo, so = o.process(bases, offsets, True, threads=8)
    eg, sg = BBDukIndexGPU.process_sharded([a] + others, bases, offsets, True)
    for name, x in eo.fields().items():
        assert np.array_equal(x, eg.fields()[name]), name
    assert so.as_dict() == sg.as_dict()
    # and the peer-copy transport across devices
    import os
    os.environ["BBDUK_B200_REPLICATE"] = "peer"
    try:
        p = a.replicate([1])[0]
    finally:
        del os.environ["BBDUK_B200_REPLICATE"]
    assert p.transport == "peer"
    with torch.cuda.device(1):
        kb, vb = p.dump_table()
    assert np.array_equal(ka, kb) and np.array_equal(va, vb)
