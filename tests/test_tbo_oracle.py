"""CPU test of the trim-by-overlap oracle (oracle/tbo_oracle.c) against an independently written Python restatement
of jgi/BBMergeOverlapper.java:411-621, :785-836 that works from mismatch COUNTS and a table of float partial sums
(the formulation the CUDA kernel uses) instead of the reference's base-by-base float accumulation."""
import numpy as np
import pytest

from bbtools_b200 import synth
from oracle import tbo

F32 = np.float32
COMP, PROB_ERROR = tbo.tables()


def partial_sums(incr, n):
    """T[c] = incr added c times in single precision, as `bad+=bIncr` does"""
    t = np.zeros(n + 1, F32)
    for c in range(1, n + 1):
        t[c] = F32(t[c - 1] + F32(incr))
    return t


def counts(a, b, insert):
    alen, blen = len(a), len(b)
    istart = 0 if insert <= blen else insert - blen
    jstart = 0 if insert >= blen else blen - insert
    ov = min(alen - istart, blen - jstart, insert)
    x, y = a[istart:istart + ov], b[jstart:jstart + ov]
    eq = x == y
    return int(np.count_nonzero(~eq)), int(np.count_nonzero(eq & (x != ord("N")))), ov


def py_find_best_ratio(a, b, mo0, mo, min_insert, max_ratio, offset, T):
    best = F32(F32(max_ratio) + F32(0.0001))
    halfmax = F32(F32(max_ratio) * F32(0.5))
    for insert in range(len(a) + len(b) - mo, min_insert - 1, -1):
        nbad, ngood, ov = counts(a, b, insert)
        badlimit = F32(F32(best * F32(ov)) + F32(20))
        # the reference stops adding once bad > badlimit; T is increasing, so "never exceeded" <=> T[nbad] <= badlimit
        if T[nbad] <= badlimit:
            bad, good = T[nbad], T[ngood]
            if bad == 0 and good > mo0 and good < mo:
                return F32(100)
            ratio = F32(F32(bad + F32(offset)) / F32(ov))
            if ratio < best:
                best = ratio
                if good >= mo and ratio < halfmax:
                    return best
    return best


def py_mate(a, b, p, T):
    mo = max(4, p.min_overlap0, p.min_overlap)
    mo0 = sorted((4, p.min_overlap0, mo))[1]
    min_len = min(len(a), len(b))
    max_ratio = F32(p.max_ratio)
    margin, offset = F32(p.ratio_margin), F32(p.ratio_offset)
    x = py_find_best_ratio(a, b, mo0, mo, p.min_insert, max_ratio, offset, T)
    if x > max_ratio:
        return -1, False
    max_ratio = min(max_ratio, x)
    margin2 = F32(F32(margin + offset) / F32(min_len))
    best_insert, best_ratio, second, ambig = -1, F32(1), F32(1), False
    for insert in range(len(a) + len(b) - mo0, p.min_insert0 - 1, -1):
        nbad, ngood, ov = counts(a, b, insert)
        badlimit = F32(F32(F32(F32(1.2) * F32(F32(min(best_ratio, max_ratio) * margin) * F32(ov))) + F32(1)) + F32(20))
        if not T[nbad] <= badlimit:
            continue
        bad, good = T[nbad], T[ngood]
        if bad == 0 and good > mo0 and good < mo:
            return -1, True
        ratio = F32(F32(bad + offset) / F32(ov))
        if ratio < F32(best_ratio * margin):
            ambig = bool(F32(ratio * margin) >= best_ratio or good < mo)
            if ratio < best_ratio:
                second, best_insert, best_ratio = best_ratio, insert, ratio
            elif ratio < second:
                second = ratio
            if (ambig and best_ratio < margin2) or second < F32(p.min_second_ratio):
                return -1, True
    if not ambig and best_ratio > max_ratio:
        best_insert = -1
    return best_insert, ambig


def py_tbo(bases, quals, offsets, lo, hi, flags, p):
    hi = hi.copy()
    T = partial_sums(p.b_incr, 700)
    ins, amb = [], []
    trimmed = [0, 0]
    for u in range(0, len(offsets) - 1, 2):
        ins.append(-1)
        amb.append(0)
        if flags[u] & 2:
            continue
        r = []
        ee = F32(0)
        for i in (u, u + 1):
            s = bases[offsets[i] + lo[i]:offsets[i] + hi[i]]
            r.append(s)
            if quals is not None:
                q = quals[offsets[i] + lo[i]:offsets[i] + hi[i]].astype(np.int64) - p.qual_offset
                e = F32(0)
                for b_, q_ in zip(s, q):
                    if chr(b_) in "ACGTUacgtu":
                        e = F32(e + PROB_ERROR[q_])
                ee = max(ee, e)
        if not ee < F32(p.mee_filter):
            continue
        a, b = r[0], COMP[r[1][::-1] & 127]
        best, ambig = py_mate(a, b, p, T)
        if best < p.min_insert:
            best = -1
        ins[-1], amb[-1] = best, int(ambig)
        if best > 0 and not ambig:
            for i, s in ((u, r[0]), (u + 1, r[1])):
                if best < len(s):
                    hi[i] = lo[i] + best
                    trimmed[0] += 1
                    trimmed[1] += len(s) - best
    return hi, np.array(ins, np.int32), np.array(amb, np.uint8), np.array(trimmed, np.int64)


def small_pairs(n, seed, L=60):
    """pairs with short inserts so that overlaps exist, plus noise, N's and unequal lengths after a fake ktrim"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = []
    for _ in range(n):
        ins = int(rng.integers(20, 2 * L + 30))
        frag = acgt[rng.integers(0, 4, max(ins, 1))]
        tail = acgt[rng.integers(0, 4, L)]
        r1 = np.concatenate([frag, tail])[:L].copy()
        r2 = np.concatenate([COMP[frag[::-1]], acgt[rng.integers(0, 4, L)]])[:L].copy()
        for r in (r1, r2):
            for _ in range(int(rng.integers(0, 4))):
                r[int(rng.integers(0, L))] = acgt[int(rng.integers(0, 4))]
            if rng.random() < 0.2:
                r[int(rng.integers(0, L))] = ord("N")
            if rng.random() < 0.05:
                r[int(rng.integers(0, L))] = rng.choice(np.frombuffer(b"acgtnRY", np.uint8))
        seqs += [r1, r2]
    bases = np.concatenate(seqs)
    offsets = np.arange(0, (2 * n + 1) * L, L, dtype=np.int64)
    lo = np.zeros(2 * n, np.int32)
    hi = np.full(2 * n, L, np.int32)
    cut = rng.random(2 * n) < 0.3
    hi[cut] -= rng.integers(1, 25, int(cut.sum())).astype(np.int32)
    flags = np.zeros(2 * n, np.uint8)
    rem = rng.random(n) < 0.05
    flags[0::2][rem] = 2
    flags[1::2][rem] = 2
    quals = (33 + rng.integers(30, 42, 2 * n * L)).astype(np.uint8).reshape(2 * n, L)
    bad_reads = rng.random(2 * n) < 0.3
    quals[bad_reads] = (33 + rng.integers(2, 16, (int(bad_reads.sum()), L))).astype(np.uint8)
    quals = quals.reshape(-1)
    return bases, quals, offsets, lo, hi, flags


@pytest.mark.parametrize("strict,with_quals,seed", [(True, False, 1), (True, True, 2), (False, False, 3), (False, True, 4)])
def test_oracle_matches_python_restatement(strict, with_quals, seed):
    bases, quals, offsets, lo, hi, flags = small_pairs(250, seed)
    p = tbo.default_params(strict)
    if with_quals and strict:
        p.mee_filter = 2.5  # make the expected-error guard bite on some pairs
    q = quals if with_quals else None
    got = tbo.process(bases, q, offsets, lo, hi, flags, p)
    want = py_tbo(bases, q, offsets, lo, hi, flags, p)
    for g, w, name in zip(got, want, ("hi", "insert", "ambig", "stats")):
        assert np.array_equal(g, w), name
    assert got[3][0] > 8  # overlaps were found and trimmed


def test_cfg2_reads_are_trimmed_to_their_insert():
    """on the cfg-2 synthetic pairs, short inserts are recovered: the trimmed length equals the insert size"""
    n = 400
    bases, offsets = synth.paired_adapter_reads(n, seed=5)
    ins_true = synth.insert_sizes(5, np.arange(n, dtype=np.uint64))
    L = np.diff(offsets).astype(np.int32)
    hi, ins, amb, st = tbo.process(bases, None, offsets, np.zeros(2 * n, np.int32), L, np.zeros(2 * n, np.uint8))
    found = (ins > 0) & (amb == 0)
    assert found.sum() > 60
    assert np.all(ins[found] == ins_true[found])
    assert not np.any(found & (ins_true >= 300))
