"""CPU tests: the C oracle against (a) SURVEY.md Appendix-B key counts, (b) the independent closed-form
restatement, (c) hand-checked cases of the reference's quirks. No GPU."""
import numpy as np
import pytest

from bbtools_b200 import make_cfg, synth
from bbtools_b200.fasta import pack
from oracle import closed_form as cf
from oracle.oracle import Oracle


def make_oracle(adapters, **kw):
    _, b, off = adapters
    o = Oracle(make_cfg(**kw))
    o.add_ref(b, off)
    o.finalize()
    return o


@pytest.mark.parametrize("kw,want", [
    (dict(k=23, ktrim_right=1), 2728),                       # cfg 1: hdist 0, maskmiddle on
    (dict(k=23, mink=11, hdist=1, ktrim_right=1), 217135),   # cfg 2: 186,778 full + 30,357 short
    (dict(k=27, hdist=2, ktrim_right=1), 8441672),           # cfg 4 on adapters.fa
])
def test_appendix_b_key_counts(adapters, kw, want):
    _, b, off = adapters
    o = Oracle(make_cfg(**kw))
    o.add_ref(b, off)
    assert o.finalize() == want
    keys, vals = o.dump_table()
    assert len(keys) == want and vals.min() >= 1 and vals.max() <= 158


def test_short_key_split(adapters):
    o = make_oracle(adapters, k=23, mink=11, hdist=1, ktrim_right=1)
    keys, _ = o.dump_table()
    full = int((keys >> np.uint64(46)).astype(np.int64).__eq__(1).sum())
    assert full == 186778 and len(keys) - full == 30357


def test_derived_constants_quirks():
    # minlen2 is computed before useShortKmers switches maskMiddle off (jgi/BBDuk.java:836 vs :849-856)
    d = Oracle(make_cfg(k=23, mink=11, hdist=1, ktrim_right=1)).derived()
    assert d["minlen2"] == 11 and d["maskMiddle"] == 0 and d["middleMask"] == -1 and d["forbidNs"] == 0
    d = Oracle(make_cfg(k=23, ktrim_right=1)).derived()
    assert d["minlen2"] == 11 and d["midMaskLen"] == 1 and d["forbidNs"] == 1
    assert d["middleMask"] == ~(3 << 22)
    d = Oracle(make_cfg(k=31)).derived()
    assert d["kfilter"] == 1 and d["minlen2"] == 15 and d["mink"] == 6
    d = Oracle(make_cfg(k=31, generation=1)).derived()
    assert d["mink"] == -1
    d = Oracle(make_cfg(k=40)).derived()  # k>31: kbig, maskmiddle off
    assert d["k"] == 31 and d["kbig"] == 40 and d["maskMiddle"] == 0 and d["minlen2"] == 31
    d = Oracle(make_cfg(k=40, ktrim_right=1)).derived()  # trimming clamps kbig
    assert d["kbig"] == 31
    with pytest.raises(ValueError):
        Oracle(make_cfg(k=23, mink=11))  # mink needs a trim mode (jgi/BBDuk.java:866)


def _small_ref():
    return ["ACGTTGCATGGATCCAGTACGATTACAGGCAT", "TTGACCAGTNNACGGATACCATGACGTTAGCAAT", "GATTACA"]


def _reads(rng, refs, n, maxlen=60, fragmax=30):
    out = []
    alpha = "ACGT" * 10 + "NnacgtRY"
    for _ in range(n):
        L = int(rng.integers(0, maxlen))
        s = [alpha[int(x)] for x in rng.integers(0, len(alpha), L)]
        if L > 6 and rng.random() < 0.7:
            r = refs[int(rng.integers(0, len(refs)))]
            a = int(rng.integers(0, len(r) - 4))
            frag = r[a:a + int(rng.integers(4, fragmax))]
            if rng.random() < 0.4:
                frag = "".join({"A": "T", "C": "G", "G": "C", "T": "A"}.get(c, c) for c in reversed(frag))
            pos = int(rng.integers(0, L))
            frag = frag[:L - pos]
            s[pos:pos + len(frag)] = list(frag)
        out.append("".join(s))
    return out


CF_CASES = [
    dict(k=11, hdist=0, mm=True),
    dict(k=11, hdist=1, mm=True),
    dict(k=12, hdist=0, mm=True),
    dict(k=11, hdist=0, mm=False),
    dict(k=11, mink=5, hdist=1, mm=True),
    dict(k=9, mink=4, hdist=0, mm=True),
    dict(k=11, hdist=1, mm=True, fn=True),
    dict(k=11, hdist=0, mm=True, rcomp=False),
]


@pytest.mark.parametrize("case", CF_CASES)
def test_closed_form_agrees_with_c_oracle(case):
    rng = np.random.default_rng(11)
    refs = _small_ref()
    reads = _reads(rng, refs, 250)
    d = cf.Derived(k=case["k"], mink=case.get("mink", -1), hdist=case["hdist"], mm=case["mm"],
                   rcomp=case.get("rcomp", True), fn=case.get("fn", False))
    table = cf.build_table(d, refs)
    rb, ro = pack([r.encode() for r in refs])
    qb, qo = pack([r.encode() for r in reads])
    common = dict(k=case["k"], mink=case.get("mink", -1), hdist=case["hdist"], mask_middle=int(case["mm"]),
                  rcomp=int(case.get("rcomp", True)), forbid_ns=int(case.get("fn", False)), min_read_length=0)

    def run(**mode):
        o = Oracle(make_cfg(**common, **mode))
        o.add_ref(rb, ro)
        n = o.finalize()
        return o, n

    # table: same key -> id map
    o, n = run(ktrim_right=1)
    keys, vals = o.dump_table()
    assert n == len(table)
    assert dict(zip((int(x) for x in keys), (int(x) for x in vals))) == table
    # ktrim=r
    out, _ = o.process(qb, qo, False)
    for i, s in enumerate(reads):
        hi, id0 = cf.ktrim_right(d, table, s)
        assert (out.hi[i], out.id0[i], out.lo[i]) == (hi, id0, 0), (i, s)
    # ktrim=l
    o, _ = run(ktrim_left=1)
    out, _ = o.process(qb, qo, False)
    for i, s in enumerate(reads):
        lo, hi, id0 = cf.ktrim_left(d, table, s)
        assert (out.lo[i], out.hi[i], out.id0[i]) == (lo, hi, id0), (i, s)
    # kmask
    o, _ = run(ktrim_n=1)
    out, _ = o.process(qb, qo, False, want_mask=True)
    for i, s in enumerate(reads):
        bits, id0 = cf.kmask_bits(d, table, s)
        w0 = int(out.mask_off[i])
        got = {j for j in range(len(s)) if (int(out.maskbits[w0 + (j >> 5)]) >> (j & 31)) & 1}
        assert got == bits and out.id0[i] == id0 and out.count[i] == len(bits), (i, s)
    # kfilter (no short k-mers allowed in this mode)
    if "mink" not in case:
        for mb in (0, 2):
            o, _ = run(max_bad_kmers=mb)
            out, _ = o.process(qb, qo, False)
            for i, s in enumerate(reads):
                cnt, cid = cf.kfilter_count(d, table, s, mb)
                assert (out.count[i], out.id0[i]) == (cnt, cid), (i, s)
                assert bool(out.flags[i] & 1) == (cnt > mb)


def test_trim_rule_keeps_one_base(adapters, adapter_seqs):
    # a read that is adapter from base 0: ktrim=r would trim to 0, TrimRead keeps 1 base
    # (shared/TrimRead.java:322-326); then minlen=10 discards it
    o = make_oracle(adapters, k=23, mink=11, hdist=1, ktrim_right=1)
    read = adapter_seqs[0][:60]
    b, off = pack([read.encode()])
    out, st = o.process(b, off, False)
    assert (out.lo[0], out.hi[0]) == (0, 1) and out.flags[0] & 1 and out.flags[0] & 2 and out.id0[0] == 1
    assert st.bases_ktrimmed == 60 and st.reads_ktrimmed == 1 and st.reads_out == 0
    # ktrim=l on the same read: left trim of everything flips into "keep first base"
    o = make_oracle(adapters, k=23, mink=11, hdist=1, ktrim_left=1)
    out, _ = o.process(b, off, False)
    assert (out.lo[0], out.hi[0]) == (0, 1)


def test_pair_logic_tpe_and_rieb(adapters):
    o = make_oracle(adapters, k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
    b, off = synth.paired_adapter_reads(4000, seed=5)
    out, st = o.process(b, off, True)
    hi = out.hi.reshape(-1, 2)
    fl = out.flags.reshape(-1, 2)
    kept = (fl[:, 0] & 2) == 0
    assert (hi[kept, 0] == hi[kept, 1]).mean() > 0.999  # tpe equalises whenever something was trimmed
    assert ((fl[:, 0] & 2) == (fl[:, 1] & 2)).all()      # removal is per pair
    assert st.reads_in == 8000 and st.reads_out == 2 * int(kept.sum())
    assert ((fl & 8) != 0).any()


def test_threads_do_not_change_results(adapters):
    o = make_oracle(adapters, k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
    b, off = synth.paired_adapter_reads(3000, seed=9)
    a, sa = o.process(b, off, True, threads=1)
    c, sc = o.process(b, off, True, threads=5)
    for name, x in a.fields().items():
        assert np.array_equal(x, c.fields()[name]), name
    assert sa.as_dict() == sc.as_dict()


@pytest.mark.parametrize("case", [dict(k=11, hdist=1, mm=False), dict(k=11, hdist=1, mm=True), dict(k=11, mink=5, hdist=1, mm=True),
                                  dict(k=11, hdist=1, mm=True, fn=True), dict(k=11, hdist=0, mm=False), dict(k=11, hdist=1, mm=False, rcomp=False)])
def test_undefined_bases_inside_reference_fragments(case):
    """The readings the GPU's undefined-base pre-pass relies on (forward k-mer: undefined = A, reverse k-mer: undefined =
    complement of T, jgi/BBDuk.java:3882-3888; forbidNs: reset), pinned on the C oracle by the independent closed form:
    every position of reference fragments of both strands replaced by N / IUPAC / lower-case n, alone and in pairs."""
    rng = np.random.default_rng(5)
    refs = _small_ref()
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    reads = []
    for r in refs[:2]:
        for frag in (r[:24], "".join(comp.get(c, c) for c in reversed(r[:24]))):
            for i in range(len(frag)):
                for gap, sym in ((0, "N"), (3, "R"), (7, "n")):
                    s = list(frag)
                    s[i] = "N"
                    if gap and i + gap < len(s):
                        s[i + gap] = sym
                    left = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, int(rng.integers(0, 12))))
                    right = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, int(rng.integers(0, 8))))
                    reads.append(left + "".join(s) + right)
    d = cf.Derived(k=case["k"], mink=case.get("mink", -1), hdist=case["hdist"], mm=case["mm"], rcomp=case.get("rcomp", True),
                   fn=case.get("fn", False))
    table = cf.build_table(d, refs)
    rb, ro = pack([r.encode() for r in refs])
    qb, qo = pack([r.encode() for r in reads])
    common = dict(k=case["k"], mink=case.get("mink", -1), hdist=case["hdist"], mask_middle=int(case["mm"]),
                  rcomp=int(case.get("rcomp", True)), forbid_ns=int(case.get("fn", False)), min_read_length=0)
    o = Oracle(make_cfg(**common, ktrim_right=1))
    o.add_ref(rb, ro)
    o.finalize()
    out, _ = o.process(qb, qo, False)
    hits = 0
    for i, s in enumerate(reads):
        hi, id0 = cf.ktrim_right(d, table, s)
        assert (out.hi[i], out.id0[i], out.lo[i]) == (hi, id0, 0), (i, s)
        hits += id0 > 0
    o = Oracle(make_cfg(**common, ktrim_left=1))
    o.add_ref(rb, ro)
    o.finalize()
    out, _ = o.process(qb, qo, False)
    for i, s in enumerate(reads):
        lo, hi, id0 = cf.ktrim_left(d, table, s)
        assert (out.lo[i], out.hi[i], out.id0[i]) == (lo, hi, id0), (i, s)
    if not case.get("fn") and case["hdist"] > 0:
        assert hits > len(reads) // 3  # the cases do exercise hits through windows with undefined bases


# ---- the modes and query options served only by the generic GPU kernel, against the closed form --------------------
def _long_ref():
    rng = np.random.default_rng(23)
    return ["".join("ACGT"[int(x)] for x in rng.integers(0, 4, n)) for n in (70, 55, 90, 64)] + _small_ref()[1:2]


def _cf_pair(refs, n_reads, maxlen, seed, gen=0, trimming=True, fragmax=30, **case):
    rng = np.random.default_rng(seed)
    reads = _reads(rng, refs, n_reads, maxlen, fragmax)
    d = cf.Derived(k=case["k"], mink=case.get("mink", -1), hdist=case.get("hdist", 0), mm=case.get("mm", True),
                   rcomp=case.get("rcomp", True), fn=case.get("fn", False), generation=gen, qhdist=case.get("qhdist", 0),
                   qskip=case.get("qskip", 1), speed=case.get("speed", 0), trimming=trimming)
    table = cf.build_table(d, refs)
    rb, ro = pack([r.encode() for r in refs])
    qb, qo = pack([r.encode() for r in reads])
    common = dict(k=case["k"], mink=case.get("mink", -1), hdist=case.get("hdist", 0), mask_middle=int(case.get("mm", True)),
                  rcomp=int(case.get("rcomp", True)), forbid_ns=int(case.get("fn", False)), min_read_length=0,
                  qhdist=case.get("qhdist", 0), qskip=case.get("qskip", 1), speed=case.get("speed", 0), generation=gen)

    def run(**mode):
        o = Oracle(make_cfg(**common, **mode))
        o.add_ref(rb, ro)
        n = o.finalize()
        keys, vals = o.dump_table()
        assert n == len(table) and dict(zip((int(x) for x in keys), (int(x) for x in vals))) == table
        return o.process(qb, qo, False)[0]

    return d, table, reads, run


QUERY_CASES = [
    dict(k=11, hdist=0, mm=True, qhdist=1),
    dict(k=11, hdist=0, mm=False, qhdist=1, rcomp=False),
    dict(k=9, mink=5, hdist=0, qhdist=1),
    dict(k=11, hdist=1, qskip=3),
    dict(k=11, hdist=1, speed=5),
    dict(k=11, hdist=1, speed=5, gen=1),
    dict(k=11, hdist=0, mm=False, speed=12, gen=1),
    dict(k=11, hdist=1, speed=9, fn=True),
]


@pytest.mark.parametrize("case", QUERY_CASES, ids=lambda c: ",".join(f"{a}={b}" for a, b in c.items()))
def test_closed_form_query_options(case):
    """qhdist (the reference's symbol-major, slot-minor depth-first order, first hit wins), qskip and both speed= rules,
    through ktrim=r and kfilter."""
    case = dict(case)
    gen = case.pop("gen", 0)
    d, table, reads, run = _cf_pair(_small_ref(), 160, 60, 31, gen=gen, **case)
    out = run(ktrim_right=1)
    for i, s in enumerate(reads):
        L = len(s)
        want_hi, want_id = L, -1
        if L >= max(1, min(d.k, d.mink) if d.usk else d.k) and table:
            hits = [(j, cf.probe(d, table, s, j)) for j in range(L)]
            hits = [(j, h) for j, h in hits if h is not None and h > 0]
            if hits:
                want_id = hits[0][1]
                _, want_hi = cf.trim_by_amount(L, 0, L - (min(j - d.k + 1 for j, _ in hits) - 1) - 1)
            elif d.usk:
                found = [(L - n, cf.tail_probe(d, table, s[L - n:], L - n)) for n in range(d.mink, min(d.k - 1, L) + 1)]
                found = [(j, h) for j, h in found if h > 0]
                if found:
                    want_id = found[0][1]
                    _, want_hi = cf.trim_by_amount(L, 0, L - (found[-1][0] - 1) - 1)
        assert (out.hi[i], out.id0[i]) == (want_hi, want_id), (i, s)
    if "mink" not in case:
        out = run()
        for i, s in enumerate(reads):
            cnt, cid = 0, -1
            if len(s) >= d.k and table:
                for j in range(len(s)):
                    h = cf.probe(d, table, s, j)
                    if h is not None and h > 0:
                        cnt, cid = 1, h
                        break
            assert (out.count[i], out.id0[i]) == (cnt, cid), (i, s)


@pytest.mark.parametrize("case", [dict(k=11, hdist=1), dict(k=11, mink=5, hdist=1), dict(k=9, mink=4, hdist=0), dict(k=11, hdist=0, fn=True),
                                  dict(k=11, hdist=1, mm=False, rcomp=False), dict(k=11, hdist=0, qhdist=1)],
                         ids=lambda c: ",".join(f"{a}={b}" for a, b in c.items()))
def test_closed_form_ktrim_tips(case):
    """ktrimTips: two passes split at the original middle, the left one over what the right one left; windows that straddle
    the start of a pass read the unseen bases as zero bits (jgi/BBDuk.java:3686-3858)."""
    d, table, reads, run = _cf_pair(_small_ref(), 300, 70, 37, **case)
    out = run(ktrim_left=1, ktrim_right=1)
    nz = 0
    for i, s in enumerate(reads):
        lo, hi, id_r, id_l, x = cf.ktrim_tips(d, table, s)
        ids = [j for j in (id_r, id_l) if j > 0] + [-1, -1]
        assert (out.lo[i], out.hi[i], out.id0[i], out.id0b[i], out.count[i]) == (lo, hi, ids[0], ids[1], x), (i, s)
        nz += x > 0
    assert nz > 30
    out = run(ktrim_left=1, ktrim_right=1, restrict_left=25, restrict_right=20)
    for i, s in enumerate(reads):
        lo, hi, id_r, id_l, x = cf.ktrim_tips(d, table, s, 25, 20)
        assert (out.lo[i], out.hi[i], out.count[i]) == (lo, hi, x), (i, s)


@pytest.mark.parametrize("case", [dict(k=11, hdist=1), dict(k=11, mink=5, hdist=1), dict(k=9, mink=4, hdist=0, fn=True)],
                         ids=lambda c: ",".join(f"{a}={b}" for a, b in c.items()))
def test_closed_form_ksplit(case):
    d, table, reads, run = _cf_pair(_small_ref(), 300, 70, 41, **case)
    out = run(ksplit=1)
    n_split = 0
    for i, s in enumerate(reads):
        lo, hi, split, at, id0 = cf.ksplit(d, table, s)
        assert (out.lo[i], out.hi[i], bool(out.flags[i] & 16), out.id0[i]) == (lo, hi, split, id0), (i, s)
        if split:
            assert out.count[i] == at
            n_split += 1
    assert n_split > 10


@pytest.mark.parametrize("case", [dict(k=11, hdist=0), dict(k=11, hdist=1, mm=False), dict(k=11, hdist=0, fn=True, qskip=2)],
                         ids=lambda c: ",".join(f"{a}={b}" for a, b in c.items()))
def test_closed_form_covered_bases_and_best_match(case):
    d, table, reads, run = _cf_pair(_small_ref(), 250, 70, 43, trimming=False, **case)
    frac = 0.3
    out = run(min_covered_fraction=frac)
    for i, s in enumerate(reads):
        need = int(np.ceil(float(np.float32(frac) * np.float32(len(s)))))
        cnt, cid = cf.covered_bases(d, table, s, need)
        assert (out.count[i], out.id0[i], bool(out.flags[i] & 1)) == (cnt, cid, cnt >= need), (i, s)  # an empty read needs 0 bases and is discarded, as in the reference
    out = run(find_best_match=1)
    for i, s in enumerate(reads):
        bid = cf.best_match(d, table, s)
        assert (out.count[i], out.id0[i], bool(out.flags[i] & 1)) == (bid, bid, bid > 0), (i, s)


@pytest.mark.parametrize("case,mb", [(dict(k=36, hdist=0), 0), (dict(k=40, hdist=0), 2), (dict(k=33, hdist=1, fn=True), 0)],
                         ids=["k=36", "k=40,mbk=2", "k=33,hdist=1,fn"])
def test_closed_form_count_big(case, mb):
    """k > 31: runs of consecutive 31-mer hits (jgi/BBDuk.java:3596-3677)."""
    refs = _long_ref()
    d, table, reads, run = _cf_pair(refs, 400, 110, 47, trimming=False, fragmax=95, **case)
    assert d.kbig == case["k"] and d.k == 31
    out = run(max_bad_kmers=mb)
    n_hit = 0
    for i, s in enumerate(reads):
        cnt, cid = cf.count_big(d, table, s, mb)
        assert (out.count[i], out.id0[i], bool(out.flags[i] & 1)) == (cnt, cid, cnt > mb), (i, s)
        n_hit += cnt > 0
    assert n_hit > 5
