"""GPU parity of trim-by-overlap (bbduk_b200_tbo / _tbo_device) against the oracle: trimmed coordinates, TBO flags,
the insert size used, and the two counters, bit for bit -- including the float decisions (ratios, margins, ambiguity)."""
import numpy as np
import pytest

from bbtools_b200 import F_TBO, make_cfg, synth
from bbtools_b200._abi import Outputs
from oracle import tbo as otbo
from test_tbo_oracle import small_pairs

pytestmark = pytest.mark.gpu


def engine():
    from bbtools_b200.bbduk import BBDukIndexGPU
    return BBDukIndexGPU(make_cfg(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1))


def as_out(lo, hi, flags):
    o = Outputs(len(lo))
    o.lo[:], o.hi[:], o.flags[:] = lo, hi, flags
    return o


def check(g, bases, quals, offsets, lo, hi, flags, strict=True, mee=0.0):
    p = otbo.default_params(strict)
    if mee:
        p.mee_filter = mee
    whi, wins, wamb, wst = otbo.process(bases, quals, offsets, lo, hi, flags, p)
    out = as_out(lo, hi, flags)
    ins, st = g.tbo(bases, quals, offsets, out, g.tbo_cfg(strict_overlap=int(strict), mee_filter=mee))
    assert np.array_equal(out.hi, whi), f"{np.count_nonzero(out.hi != whi)} trimmed lengths differ"
    want_ins = np.where(wamb == 1, -2, np.where(wins > 0, wins, -1))
    assert np.array_equal(ins, want_ins)
    assert np.array_equal((out.flags & F_TBO) != 0, whi != hi)
    assert np.array_equal(out.flags & ~np.uint8(F_TBO), flags)
    assert list(st) == list(wst)
    return wst


@pytest.mark.parametrize("strict,with_quals,seed", [(True, False, 11), (True, True, 12), (False, False, 13), (False, True, 14)])
def test_small_pairs(strict, with_quals, seed):
    g = engine()
    bases, quals, offsets, lo, hi, flags = small_pairs(3000, seed)
    st = check(g, bases, quals if with_quals else None, offsets, lo, hi, flags, strict, 2.5 if (with_quals and strict) else 0.0)
    assert st[0] > 100


def test_after_the_kmer_block_on_cfg2_pairs(adapters):
    """the canonical command: ktrim=r k=23 mink=11 hdist=1 tpe tbo"""
    from oracle.oracle import Oracle
    g = engine()
    _, rb, roff = adapters
    g.add_ref(rb, roff)
    g.finalize()
    o = Oracle(g.cfg)
    o.add_ref(rb, roff)
    o.finalize()
    bases, offsets = synth.paired_adapter_reads(40000, seed=7)
    quals = np.full(len(bases), 33 + 40, np.uint8)
    out, _ = g.process(bases, offsets, True)
    want, _ = o.process(bases, offsets, True)
    assert np.array_equal(out.hi, want.hi) and np.array_equal(out.flags, want.flags)
    whi, wins, wamb, wst = otbo.process(bases, quals, offsets, want.lo, want.hi, want.flags)
    ins, st = g.tbo(bases, quals, offsets, out)
    assert np.array_equal(out.hi, whi) and list(st) == list(wst)
    assert st[0] > 1000  # overlaps the k-mers missed (adapter shorter than mink at the read end, or mutated)


def test_long_ragged_and_degenerate():
    g = engine()
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = []
    for n1, n2, ins in ((700, 650, 500), (1008, 1008, 900), (0, 50, 0), (50, 0, 0), (5, 5, 0), (300, 40, 120), (16, 16, 16)):
        frag = acgt[rng.integers(0, 4, max(ins, 1))]
        r1 = np.concatenate([frag, acgt[rng.integers(0, 4, 1100)]])[:n1]
        comp = otbo.tables()[0]
        r2 = np.concatenate([comp[frag[::-1]], acgt[rng.integers(0, 4, 1100)]])[:n2]
        seqs += [r1, r2]
    bases = np.concatenate(seqs).astype(np.uint8)
    offsets = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=offsets[1:])
    L = np.diff(offsets).astype(np.int32)
    z = np.zeros(len(L), np.int32)
    st = check(g, bases, None, offsets, z, L, np.zeros(len(L), np.uint8))
    assert st[0] >= 2
    # a read beyond the device limit is an error, not a silent CPU path
    big = np.concatenate([bases, acgt[rng.integers(0, 4, 2100)]]).astype(np.uint8)
    off2 = np.concatenate([offsets, [offsets[-1] + 1050, offsets[-1] + 2100]])
    L2 = np.diff(off2).astype(np.int32)
    with pytest.raises(RuntimeError, match="1008"):
        g.tbo(big, None, off2, as_out(np.zeros(len(L2), np.int32), L2, np.zeros(len(L2), np.uint8)))


def test_device_entry_point():
    import torch
    g = engine()
    bases, offsets = synth.paired_adapter_reads(30000, seed=9)
    n = len(offsets) - 1
    L = np.diff(offsets).astype(np.int32)
    lo, flags = np.zeros(n, np.int32), np.zeros(n, np.uint8)
    whi, wins, wamb, wst = otbo.process(bases, None, offsets, lo, L, flags)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_hi, d_flags = d(L), d(flags)
    d_ins = torch.empty(n // 2, dtype=torch.int32, device="cuda")
    d_st = torch.zeros(2, dtype=torch.int64, device="cuda")
    g.tbo_device(d(bases), None, d(offsets.astype(np.int32)), n, 150, d(lo), d_hi, d_flags, d_ins, d_st)
    torch.cuda.synchronize()
    assert np.array_equal(d_hi.cpu().numpy(), whi)
    assert d_st.cpu().tolist() == list(wst)


@pytest.mark.parametrize("strict", [True, False])
def test_coincident_ns_and_borderline_mismatch_rates(strict):
    """true overlaps full of Ns at the SAME fragment positions in both mates (N against N is no mismatch, the screen must not
    count it), mismatch rates straddling maxRatio, tandem repeats"""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    comp, _ = otbo.tables()
    seqs = []
    for i in range(3000):
        n1, n2 = (int(x) for x in rng.integers(60, 260, 2))
        ins = int(rng.integers(20, n1 + n2))
        if i % 4 == 0:
            unit = acgt[rng.integers(0, 4, int(rng.integers(1, 9)))]
            frag = np.tile(unit, ins // len(unit) + 1)[:ins].copy()
        else:
            frag = acgt[rng.integers(0, 4, ins)].copy()
        if i % 3 == 0:
            frag[rng.random(ins) < (0.05 + 0.3 * rng.random())] = ord("N")
        r1 = np.concatenate([frag, acgt[rng.integers(0, 4, 300)]])[:n1].copy()
        r2 = np.concatenate([comp[frag[::-1]], acgt[rng.integers(0, 4, 300)]])[:n2].copy()
        rate = rng.random() * 0.2
        for r in (r1, r2):
            hit = rng.random(len(r)) < rate / 2
            r[hit] = acgt[rng.integers(0, 4, int(hit.sum()))]
        seqs += [r1, r2]
    bases = np.concatenate(seqs).astype(np.uint8)
    offsets = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(x) for x in seqs], out=offsets[1:])
    L = np.diff(offsets).astype(np.int32)
    st = check(engine(), bases, None, offsets, np.zeros(len(L), np.int32), L, np.zeros(len(L), np.uint8), strict)
    assert st[0] > 500
