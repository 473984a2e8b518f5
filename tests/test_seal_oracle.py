"""CPU tests of the Seal oracle (oracle/seal_oracle.c): hand-computed cases and an independent closed-form Python
restatement (oracle/seal.py:ClosedFormSeal) over randomized references, reads and option sets.
Reference: jgi/Seal.java:1760-1946 (loader), :2186-2276, :2386-2606, :2654-2708, :2864-2907 (matching)."""
import numpy as np
import pytest

from oracle import seal as S


def pack(strs):
    off = np.zeros(len(strs) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in strs])
    return np.frombuffer("".join(strs).encode(), np.uint8).copy(), off


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), n))


def make_case(seed, n_refs=6, ref_len=300, n_frag=60, read_len=100, paired=True, k=31, n_rate=0.01):
    """References that share stretches (multi-id k-mers) + reads sampled from them with errors, Ns and junk."""
    rng = np.random.default_rng(seed)
    shared = rand_seq(rng, 80)
    refs = []
    for r in range(n_refs):
        s = rand_seq(rng, ref_len)
        if r % 2 == 0:
            p = int(rng.integers(0, ref_len - 80))
            s = s[:p] + shared + s[p + 80:]
        if r == 3:
            s = s[:50] + "N" + s[51:120] + "R" + s[121:]
        refs.append(s)
    refs.append(refs[1][:150])  # a reference contained in another
    refs.append("ACGT" * 3)     # shorter than k for k=31
    comp = str.maketrans("ACGTN", "TGCAN")

    def sample():
        r = rng.integers(0, 10)
        if r == 0:
            s = rand_seq(rng, read_len)
        else:
            src = refs[int(rng.integers(0, len(refs) - 1))]
            L = int(min(len(src), rng.integers(k - 2, read_len + 1)))
            p = int(rng.integers(0, len(src) - L + 1))
            s = src[p:p + L]
            if rng.integers(0, 4) == 0:  # chimera of two references
                src2 = refs[int(rng.integers(0, n_refs))]
                s = s[:L // 2] + src2[:L - L // 2]
            if rng.integers(0, 2):
                s = s.translate(comp)[::-1]
        s = list(s)
        for i in range(len(s)):
            u = rng.random()
            if u < n_rate:
                s[i] = "N"
            elif u < n_rate * 1.5:
                s[i] = rng.choice(list("acgtRYn."))
            elif u < n_rate * 3:
                s[i] = rng.choice(list("ACGT"))
        return "".join(s)

    reads = [sample() for _ in range(n_frag * (2 if paired else 1))]
    if len(reads) > 4:
        reads[2] = ""
        reads[3] = "ACG"
    return refs, reads


OPTION_SETS = [
    dict(),
    dict(k=31, ambig_mode=S.AMBIG_ALL),
    dict(k=25, ambig_mode=S.AMBIG_FIRST, clearzone=3),
    dict(k=27, ambig_mode=S.AMBIG_TOSS, mask_middle=0),
    dict(k=21, hdist=1, mask_middle=1, mid_mask_len=3),
    dict(k=13, hdist=2, mask_middle=0, ambig_mode=S.AMBIG_ALL),
    dict(k=20, hdist=1, forbid_ns=1, clearzone_fraction=0.1),
    dict(k=31, keep_pairs_together=0),
    dict(k=24, keep_pairs_together=0, ambig_mode=S.AMBIG_ALL, clearzone=5, min_kmer_hits=3),
    dict(k=31, match_mode=S.MATCH_FIRST),
    dict(k=31, match_mode=S.MATCH_UNIQUE, ambig_mode=S.AMBIG_ALL),
    dict(k=23, restrict_left=60),
    dict(k=23, restrict_right=50, keep_pairs_together=0),
    dict(k=31, qskip=3),
    dict(k=31, speed=6),
    dict(k=31, rskip=4),
    dict(k=31, min_kmer_fraction=0.3),
    dict(k=19, rcomp=0, clearzone_fraction=0.05, clearzone=2),
    dict(k=5, mask_middle=0, ambig_mode=S.AMBIG_ALL, ids_stride=8),
    dict(k=31, mid_mask_len=5, hdist=1),
]


def compare(cfg, refs, reads, paired, first_id=7):
    ora = S.SealOracle(cfg)
    ora.add_ref(*pack(refs))
    v = ora.finalize()
    cf = S.ClosedFormSeal(cfg)
    cf.add_ref(refs)
    assert cf.finalize() == v
    ck, ci = cf.table_arrays()
    ok, oi = ora.table()
    assert np.array_equal(ck, ok) and np.array_equal(ci, oi)
    res, st = ora.process(*pack(reads), paired, first_id)
    units, cst, sc = cf.process(reads, paired, first_id)
    assert len(units) == len(res.n_assigned)
    stride = cfg.ids_stride
    for u, (got, sites, mx) in enumerate(units):
        assert res.n_assigned[u] == len(got), u
        assert res.first_id[u] == (got[0] if got else 0), u
        assert res.n_sites[u] == sites and res.max_hits[u] == mx, u
        want = (got + [0] * stride)[:stride]
        assert list(res.ids[u * stride:(u + 1) * stride]) == want, u
    assert st.as_dict() == cst
    for a, b in zip(ora.scaffold_counts(), sc):
        assert np.array_equal(a, b)
    return res, st


@pytest.mark.parametrize("i", range(len(OPTION_SETS)))
def test_oracle_vs_closed_form(i):
    kw = OPTION_SETS[i]
    cfg = S.make_cfg(**kw)
    small = cfg.hdist == 2
    refs, reads = make_case(100 + i, n_refs=3 if small else 6, ref_len=120 if small else 300, n_frag=30, k=cfg.k)
    res, st = compare(cfg, refs, reads, paired=True)
    assert st.reads_matched > 0
    refs, reads = make_case(200 + i, n_refs=3 if small else 5, ref_len=120 if small else 300, n_frag=25, paired=False, k=cfg.k)
    compare(cfg, refs, reads, paired=False)


def test_hand_computed():
    # two references sharing their first 40 bases; k=31 no middle mask: a read over the shared part hits both 10 times
    rng = np.random.default_rng(5)
    shared = rand_seq(rng, 40)
    a, b = shared + "A" + rand_seq(rng, 59), shared + "C" + rand_seq(rng, 59)
    cfg = S.make_cfg(k=31, mask_middle=0, ambig_mode=S.AMBIG_ALL)
    ora = S.SealOracle(cfg)
    ora.add_ref(*pack([a, b]))
    stored, entries, refk = ora.finalize()
    assert refk == 140 and entries == 140 and stored == 130  # 10 shared 31-mers
    reads = [shared, a[20:80], b[:31], "N" * 50]
    res, st = ora.process(*pack(reads), False, 0)
    assert list(res.n_assigned) == [2, 1, 2, 0]
    assert list(res.max_hits) == [10, 30, 1, 0]
    assert list(res.n_sites) == [2, 1, 2, 0]
    assert list(res.first_id) == [1, 1, 1, 0]
    # a[20:80]: all 30 windows contain a[40], the first base that differs: id 1 only
    assert st.as_dict() == dict(reads_in=4, bases_in=40 + 60 + 31 + 50, reads_matched=3, bases_matched=131,
                                reads_unmatched=1, bases_unmatched=50)
    r, bs, fr, am = ora.scaffold_counts()
    assert list(r) == [0, 3, 2] and list(fr) == [0, 3, 2] and list(am) == [0, 2, 2]
    assert list(bs) == [0, 131, 71]
    # ambig=random: pair id picks finalList[numericID % sites]
    cfg = S.make_cfg(k=31, mask_middle=0)
    ora = S.SealOracle(cfg)
    ora.add_ref(*pack([a, b]))
    ora.finalize()
    for nid in range(4):
        res, _ = ora.process(*pack([shared]), False, nid)
        assert res.first_id[0] == 1 + nid % 2 and res.n_assigned[0] == 1
    # a[5:70]: 35 windows, the 5 that start at a[5..9] lie inside the shared part: 35 hits for id 1, 5 for id 2;
    # thresh = max - cz: a clear zone of 30 lets id 2 through, 29 does not
    for cz, want in ((30, [1, 2]), (29, [1, 0])):
        cfg = S.make_cfg(k=31, mask_middle=0, ambig_mode=S.AMBIG_ALL, clearzone=cz)
        ora = S.SealOracle(cfg)
        ora.add_ref(*pack([a, b]))
        ora.finalize()
        res, _ = ora.process(*pack([a[5:70]]), False, 0)
        assert res.max_hits[0] == 35 and list(res.ids[:2]) == want


def test_default_middle_mask_and_n():
    # default mm=t, k=31: the middle base is a wildcard, and with forbidn a window is probed once 15 bases follow an N
    rng = np.random.default_rng(9)
    a = rand_seq(rng, 80)
    cfg = S.make_cfg()
    ora = S.SealOracle(cfg)
    ora.add_ref(*pack([a]))
    ora.finalize()
    r = list(a[:31])
    r[15] = "ACGT"[("ACGT".index(r[15]) + 1) & 3]
    res, _ = ora.process(*pack(["".join(r)]), False, 0)
    assert res.max_hits[0] == 1
    r[14] = "ACGT"[("ACGT".index(r[14]) + 1) & 3]
    res, _ = ora.process(*pack(["".join(r)]), False, 0)
    assert res.max_hits[0] == 0
    cf = S.ClosedFormSeal(cfg)
    cf.add_ref([a])
    cf.finalize()
    read = a[:20] + "N" + a[21:70]
    res, _ = ora.process(*pack([read]), False, 0)
    units, _, _ = cf.process([read], False, 0)
    assert units[0][2] == res.max_hits[0]


def _two_refs(seed=5):
    rng = np.random.default_rng(seed)
    shared = rand_seq(rng, 40)
    return shared, shared + "A" + rand_seq(rng, 59), shared + "C" + rand_seq(rng, 59)


def _run(kw, refs, reads, paired, nid=0):
    ora = S.SealOracle(S.make_cfg(mask_middle=0, **kw))
    ora.add_ref(*pack(refs))
    ora.finalize()
    res, st = ora.process(*pack(reads), paired, nid)
    return res, st.as_dict(), ora.scaffold_counts()


def test_hand_computed_pair_rules():
    """Counts worked out by hand from the Java (k=31, mm=f; a and b share their first 40 bases = 10 k-mers)."""
    shared, a, b = _two_refs()
    r1, r2 = a[40:100], a[45:80]  # 30 and 5 windows, reference 1 only
    # kept together (jgi/Seal.java:2386-2453): one unit, readSum 2, one fragment, bases of both mates
    res, st, (rd, bs, fr, am) = _run(dict(), [a, b], [r1, r2], True)
    assert list(res.max_hits) == [35] and list(res.first_id) == [1] and list(res.n_assigned) == [1]
    assert (rd[1], bs[1], fr[1], am[1]) == (2, 95, 1, 0) and st["reads_matched"] == 2 and st["bases_matched"] == 95
    # apart (:2462-2606): the fragment goes to the mate with more hits, ties to read 1 (max1>=max2, max2>max1)
    res, st, (rd, bs, fr, am) = _run(dict(keep_pairs_together=0), [a, b], [r1, r2], True)
    assert list(res.max_hits) == [30, 5] and (rd[1], bs[1], fr[1]) == (2, 95, 1)
    res, st, (rd, bs, fr, am) = _run(dict(keep_pairs_together=0), [a, b], [r2, r1], True)
    assert list(res.max_hits) == [5, 30] and fr[1] == 1
    res, st, (rd, bs, fr, am) = _run(dict(keep_pairs_together=0), [a, b], [r1, r1], True)
    assert fr[1] == 1 and rd[1] == 2  # tie: read 1 only
    # minimum hits: below mkh a kept-together pair is UNMATCHED (:2226-2227), mates taken apart are counted nowhere (:2467, :2535)
    res, st, _ = _run(dict(min_kmer_hits=6), [a, b], [r2, r2[:34]], True)
    assert list(res.max_hits) == [9] and st["reads_matched"] == 2
    res, st, _ = _run(dict(min_kmer_hits=10), [a, b], [r2, r2[:34]], True)
    assert list(res.n_assigned) == [0] and st["reads_unmatched"] == 2 and st["bases_unmatched"] == 69
    res, st, _ = _run(dict(min_kmer_hits=5, keep_pairs_together=0), [a, b], [r2, r2[:34]], True)
    assert list(res.n_assigned) == [1, 0] and (st["reads_matched"], st["reads_unmatched"]) == (1, 0)
    # mkf: minhits = max(mkh, (int)(mkf * numKmers)) (:2223): 35 reference bases + 20 others = 25 windows, 5 of them hit
    x = r2 + rand_seq(np.random.default_rng(77), 20)
    res, st, _ = _run(dict(min_kmer_fraction=0.2), [a, b], [x], False)   # (int)(0.2f * 25) = 5 <= 5
    assert list(res.max_hits) == [5] and list(res.n_assigned) == [1]
    res, st, _ = _run(dict(min_kmer_fraction=0.24), [a, b], [x], False)  # (int)(0.24f * 25) = 6 > 5 (0.24f*25 = 6.0000001)
    assert list(res.max_hits) == [5] and list(res.n_assigned) == [0] and st["reads_unmatched"] == 1


def test_hand_computed_scan_rules():
    shared, a, b = _two_refs()
    # match=first (:2902): the scan stops at the first hit
    res, _, _ = _run(dict(match_mode=S.MATCH_FIRST, ambig_mode=S.AMBIG_ALL), [a, b], [a[:70]], False)
    assert list(res.max_hits) == [1] and list(res.n_sites) == [2]  # the first window lies in the shared part: both ids, once
    # match=unique: every hit counts until the first k-mer with ONE id (inclusive): 10 shared windows, then a[10:41]
    res, _, _ = _run(dict(match_mode=S.MATCH_UNIQUE, ambig_mode=S.AMBIG_ALL), [a, b], [a[:70]], False)
    assert list(res.max_hits) == [11] and list(res.n_sites) == [1] and list(res.first_id) == [1]
    res, _, _ = _run(dict(ambig_mode=S.AMBIG_ALL), [a, b], [a[:70]], False)
    assert list(res.max_hits) == [40] and list(res.n_sites) == [1]     # 40 windows for id 1, 10 for id 2
    # restrictleft=35: windows ending at 30..34 (:2877-2878)
    res, _, _ = _run(dict(restrict_left=35), [a, b], [a[40:100]], False)
    assert list(res.max_hits) == [5]
    # restrictright=35 on 60 bases: the registers start at base 25, len reaches k at base 55: windows ending at 55..59
    res, _, _ = _run(dict(restrict_right=35), [a, b], [a[40:100]], False)
    assert list(res.max_hits) == [5]
    # qskip=3: windows ending at 30, 33, ..., 57 (:2792)
    res, _, _ = _run(dict(qskip=3), [a, b], [a[40:100]], False)
    assert list(res.max_hits) == [10]
    # rskip=3 on a reference without undefined bases: the k-mers ending at run lengths 33, 36, ..., 99 are stored (:1794)
    ora = S.SealOracle(S.make_cfg(mask_middle=0, rskip=3))
    ora.add_ref(*pack([a]))
    assert ora.finalize() == (23, 23, 70)
    # reverse strand: the reverse complement of a read hits the same k-mers (rcomp=t), and none with rcomp=f
    comp = str.maketrans("ACGT", "TGCA")
    rc = a[40:100].translate(comp)[::-1]
    res, _, _ = _run(dict(), [a, b], [rc], False)
    assert list(res.max_hits) == [30]
    res, _, _ = _run(dict(rcomp=0), [a, b], [rc], False)
    assert list(res.max_hits) == [0]
    # ambig=first takes the smallest id, toss nothing, random the (numericID % sites)-th in first-seen order
    for mode, nid, want in ((S.AMBIG_FIRST, 0, [1]), (S.AMBIG_TOSS, 0, []), (S.AMBIG_RANDOM, 5, [2]), (S.AMBIG_ALL, 0, [1, 2])):
        res, st, _ = _run(dict(ambig_mode=mode), [a, b], [shared], False, nid)
        assert list(res.ids[:res.n_assigned[0]]) == want and res.n_sites[0] == 2
        assert st["reads_matched"] == (1 if want else 0)


@pytest.mark.parametrize("kw", [dict(), dict(keep_pairs_together=0, ambig_mode=S.AMBIG_ALL, clearzone=4), dict(ambig_mode=S.AMBIG_FIRST, k=21, hdist=1)])
def test_threads_do_not_change_results(kw):
    """sl_ora_process_mt (the reference arm of bench.py --workload seal): contiguous slices, counters summed"""
    cfg = S.make_cfg(**kw)
    refs, reads = make_case(55, n_refs=8, ref_len=400, n_frag=500, k=cfg.k, read_len=120)
    outs = []
    for threads in (1, 2, 7):
        ora = S.SealOracle(cfg)
        ora.add_ref(*pack(refs))
        ora.finalize()
        res, st = ora.process(*pack(reads), True, 99, threads=threads)
        outs.append(([x.copy() for x in res.fields().values()], st.as_dict(), [c.copy() for c in ora.scaffold_counts()]))
    for other in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(outs[0][0], other[0]))
        assert outs[0][1] == other[1]
        assert all(np.array_equal(a, b) for a, b in zip(outs[0][2], other[2]))


def test_random_flag_combinations_oracle_vs_closed_form():
    """the flag generator of tools/stress_seal_gpu.py (GPU vs oracle) turned on the oracle itself: C restatement vs the closed form"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from stress_seal_gpu import random_flags
    rng = np.random.default_rng(2024)
    for c in range(30):
        kw = random_flags(rng)
        cfg = S.make_cfg(**kw)
        small = cfg.hdist == 2
        paired = bool(rng.integers(0, 2))
        refs, reads = make_case(int(rng.integers(1, 1 << 30)), n_refs=3 if small else int(rng.integers(2, 7)), ref_len=120 if small else 250,
                                n_frag=20, read_len=int(rng.integers(40, 200)), paired=paired, k=cfg.k,
                                n_rate=float(rng.choice([0.0, 0.01, 0.05])))
        try:
            compare(cfg, refs, reads, paired=paired, first_id=int(rng.integers(0, 1 << 40)))
        except AssertionError as e:
            raise AssertionError(f"case {c}: flags {kw}, paired {paired}: {e}")
