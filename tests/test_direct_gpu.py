"""GPU parity of the HBM-resident-table path (probe_direct.cu + the stage-D epilogue) against the oracle:
references large enough that neither on-chip filter applies (BASELINE.json configs 3 and 4, scaled so the
oracle builds its table in seconds). Bit-exact on every output array, counters and per-scaffold counts."""
import numpy as np
import pytest

from bbtools_b200 import make_cfg, synth
from test_parity_gpu import assert_same, check_scaffold_counts, engines

pytestmark = pytest.mark.gpu


def reads_from_reference(ref_bases, n, L, seed, frag_min=10):
    """ragged reads: random bases with a fragment of the reference (either strand) spliced in, plus N/IUPAC/lowercase"""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    seqs = []
    for i in range(n):
        ln = int(rng.integers(0, L + 1))
        s = alpha[rng.integers(0, 4, ln)]
        if ln > frag_min and rng.random() < 0.5:
            fl = int(rng.integers(frag_min, ln + 1))
            st = int(rng.integers(0, len(ref_bases) - fl))
            frag = np.array(ref_bases[st:st + fl])
            if rng.random() < 0.5:
                frag = synth._comp_ascii(frag[::-1].copy())
            pos = int(rng.integers(0, ln - fl + 1))
            s[pos:pos + fl] = frag
            for _ in range(int(rng.integers(0, 3))):  # a few substitutions
                s[int(rng.integers(0, ln))] = alpha[int(rng.integers(0, 4))]
        if ln and rng.random() < 0.25:
            for _ in range(int(rng.integers(1, 4))):
                s[int(rng.integers(0, ln))] = rng.choice(np.frombuffer(b"NnRYacgtU", np.uint8))
        seqs.append(s)
    off = np.zeros(n + 1, np.int64)
    np.cumsum([len(x) for x in seqs], out=off[1:])
    return (np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)).astype(np.uint8), off


def uses_direct_path(g, b, off, paired):
    l0 = g.launches
    g.process(b, off, paired)
    return g.launches - l0 >= 5  # off64->32, init, starts, direct, epilogue


CASES = [
    (dict(k=31), (4, 600_000)),                                     # cfg-3 shape: kfilter, mm=t, forbidNs
    (dict(k=31, mask_middle=0, rcomp=0, skip_r1=1), (2, 700_000)),
    (dict(k=27, hdist=1, ktrim_right=1, trim_pairs_evenly=1), (40, 1000)),   # cfg-4 style neighbourhood, ktrim=r
    (dict(k=27, hdist=1, ktrim_left=1, trim_pad=1, min_len_fraction=0.3), (40, 1000)),
    (dict(k=25, hdist=1, ktrim_right=1, ktrim_exclusive=1, forbid_ns=1, require_both_bad=1), (50, 900)),
    (dict(k=21, hdist=1), (30, 1500)),                              # kfilter with hdist
]


@pytest.mark.parametrize("kw,shape", CASES, ids=lambda x: ",".join(f"{a}={b}" for a, b in x.items()) if isinstance(x, dict) else str(x))
def test_direct_modes(kw, shape):
    ref = synth.random_reference(shape[0], shape[1], seed=7)
    o, g = engines(None, ref=ref, **kw)
    b, off = reads_from_reference(ref[0], 6000, 200, seed=31)
    assert uses_direct_path(g, b[:off[64]], off[:65], False), "this table should take the direct path"
    o.process(b[:off[64]], off[:65], False)  # keep the per-scaffold counters of both sides in step
    assert_same(o, g, b, off, False)
    assert_same(o, g, b, off, True)
    cb, co = synth.contaminant_reads(20000, ref[0], seed=2, contam_pct=20)
    assert_same(o, g, cb, co, False)
    assert_same(o, g, cb, co, True)
    # edge cases: empty batch members, reads shorter than k, a read that is one long reference stretch
    seqs = [b"", b"ACGT", bytes(ref[0][:5000]), b"", bytes(ref[0][100:100 + kw["k"]]), bytes(ref[0][200:200 + kw["k"] - 1]), b"N" * 50,
            bytes(ref[0][300:400])]
    from bbtools_b200.fasta import pack
    pb, po = pack(seqs)
    assert_same(o, g, pb, po, False)
    assert_same(o, g, pb, po, True)
    check_scaffold_counts(o, g)


def test_direct_device_entry_point():
    import torch
    ref = synth.random_reference(3, 500_000, seed=7)
    o, g = engines(None, ref=ref, k=31)
    hb, ho = synth.contaminant_reads(50000, ref[0], seed=5, contam_pct=10)
    n = len(ho) - 1
    d_bases = torch.from_numpy(hb).cuda()
    d_off = torch.from_numpy(ho.astype(np.int32)).cuda()
    outs = {k: torch.empty(n, dtype=torch.int32, device="cuda") for k in ("id0", "hi", "lo", "count", "id0b")}
    outs["flags"] = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    g.set_max_read_len(150)
    g.process_device(d_bases, d_off, n, False, outs, d_stats=d_stats)
    torch.cuda.synchronize()
    eo, so = o.process(hb, ho, False, threads=8)
    for name in ("id0", "id0b", "lo", "hi", "count", "flags"):
        assert np.array_equal(outs[name].cpu().numpy(), eo.fields()[name]), name
    assert d_stats.cpu().tolist() == list(so.as_dict().values())
    assert so.reads_kfiltered > 3000
