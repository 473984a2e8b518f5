"""CPU test of the low-entropy filter oracle (oracle/entropy_oracle.c) against a Python restatement of
tracker/EntropyTracker.java (:62-118, :194-201, :657-703, :815-946) that keeps the window as an explicit deque of k-mers
and a dict of counts (same double-precision update order, which the bit-exact comparison needs), a from-scratch Shannon
entropy per window as an independent sanity check (the reference's own verify() tolerance), and hand-checked cases."""
import math
from collections import deque

import numpy as np
import pytest

from oracle import entropy as oe

CODE = {ord(c): v for c, v in zip("ACGTUacgtu", (0, 1, 2, 3, 3, 0, 1, 2, 3, 3))}


def py_average_entropy(seq, k=5, window=50):
    wk = window - k + 1
    E = [0.0] + [(i / wk) * math.log(i / wk) if False else (i * (1.0 / wk)) * math.log(i * (1.0 / wk)) for i in range(1, wk + 2)]
    mult = -1 / math.log(wk)
    counts, win = {}, deque()
    esum, total, div = 0.0, 0.0, 0
    codes = [CODE.get(b, 0) for b in seq]

    def calc():
        f = np.float32(esum * mult)
        return f if f > 0 else np.float32(0)

    def add(i):
        nonlocal esum
        if i >= k - 1:
            km = tuple(codes[i - k + 1:i + 1])
            old = counts.get(km, 0)
            counts[km] = old + 1
            win.append(km)
            esum = esum + E[old + 1] - E[old]
        if i >= window:
            km = win.popleft()
            old = counts[km]
            counts[km] = old - 1
            esum = esum + E[old - 1] - E[old]

    n = len(seq)
    i = 0
    for i in range(min(n, window)):
        add(i)
    total += float(calc())
    div += 1
    for i in range(min(n, window), n):
        add(i)
        total += float(calc())
        div += 1
        # independent check of the running sum: Shannon entropy of the window's k-mer multiset, from scratch
        c = {}
        for j in range(i - window + k, i + 1):
            km = tuple(codes[j - k + 1:j + 1])
            c[km] = c.get(km, 0) + 1
        h = -sum((v / wk) * math.log(v / wk) for v in c.values()) / math.log(wk)
        assert abs(h - float(calc())) < 1e-5
    return np.float32(total / max(1, div))


def entropy_batch(n, seed, L=120):
    """random reads, low-complexity reads (homopolymers, short tandem repeats, mixtures), N's, ragged lengths"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = []
    for _ in range(n):
        ln = int(rng.integers(0, L + 1)) if rng.random() < 0.4 else L
        kind = rng.integers(0, 6)
        if kind == 0:
            s = acgt[rng.integers(0, 4, ln)]
        elif kind == 1:
            s = np.full(ln, acgt[rng.integers(0, 4)], np.uint8)
        elif kind == 2:
            unit = acgt[rng.integers(0, 4, int(rng.integers(1, 7)))]
            s = np.tile(unit, ln // len(unit) + 1)[:ln]
        elif kind == 3:
            unit = acgt[rng.integers(0, 4, int(rng.integers(2, 5)))]
            s = np.concatenate([acgt[rng.integers(0, 4, ln)][:ln // 2], np.tile(unit, ln)[:ln - ln // 2]])
        elif kind == 4:
            s = acgt[rng.choice(4, ln, p=[0.7, 0.1, 0.1, 0.1])]
        else:
            s = acgt[rng.integers(0, 2, ln)]
        s = s.copy()
        if ln and rng.random() < 0.3:
            s[rng.integers(0, ln, int(rng.integers(1, 4)))] = rng.choice(np.frombuffer(b"NNnacgtRY", np.uint8))
        seqs.append(s)
    bases = np.concatenate(seqs).astype(np.uint8)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=offsets[1:])
    lo = np.zeros(n, np.int32)
    hi = np.diff(offsets).astype(np.int32)
    cut = rng.random(n) < 0.3
    hi[cut] = np.maximum(0, hi[cut] - rng.integers(1, 40, int(cut.sum()))).astype(np.int32)
    cutl = rng.random(n) < 0.15
    lo[cutl] = np.minimum(hi[cutl], rng.integers(1, 12, int(cutl.sum()))).astype(np.int32)
    flags = np.zeros(n, np.uint8)
    flags[rng.random(n) < 0.04] = 1
    return bases, offsets, lo, hi, flags


@pytest.mark.parametrize("k,window,seed", [(5, 50, 1), (5, 50, 2), (4, 30, 3), (3, 20, 4), (2, 12, 5)])
def test_entropy_values_match_python_restatement(k, window, seed):
    bases, offsets, lo, hi, _ = entropy_batch(150, seed)
    got = oe.values(bases, offsets, lo, hi, k, window)
    for i in range(len(lo)):
        want = py_average_entropy(bytes(bases[offsets[i] + lo[i]:offsets[i] + hi[i]]), k, window)
        assert got[i] == want, (i, got[i], want)
    assert got.min() == 0 and got.max() > 0.8


def test_known_answers():
    def e(seq, k=5, window=50):
        b = np.frombuffer(seq.encode(), np.uint8)
        return float(oe.values(b, np.array([0, len(b)], np.int64), np.zeros(1, np.int32), np.array([len(b)], np.int32), k, window)[0])
    assert e("A" * 100) == 0.0                       # one k-mer in every window
    assert e("") == 0.0
    assert e("ACGT") == 0.0                          # shorter than k: no k-mer at all
    # 50 bases with 46 distinct 5-mers: entropy 1 (a de Bruijn-like stretch), so the single window averages 1
    import itertools
    seen, s = set(), "AAAAA"
    seen.add(s)
    while len(s) < 50:
        for c in "CGTA":
            if s[-4:] + c not in seen:
                seen.add(s[-4:] + c)
                s += c
                break
        else:
            raise AssertionError
    assert abs(e(s) - 1.0) < 1e-6
    # a dinucleotide repeat has 2 distinct 5-mers, 23 copies each: H = ln 2 / ln 46
    assert abs(e("AC" * 25) - math.log(2) / math.log(46)) < 1e-6


@pytest.mark.parametrize("case", [dict(cutoff=0.5), dict(cutoff=0.7, rieb=False), dict(cutoff=0.3, tf1=True), dict(cutoff=0.9, k=4, window=30),
                                  dict(cutoff=0.5, high_pass=False), dict(cutoff=-1.0)])
def test_filter_block_matches_python(case):
    paired = True
    bases, offsets, lo, hi, flags = entropy_batch(400, 11)
    flags[0::2][np.arange(200) % 17 == 0] = 2
    flags[1::2][np.arange(200) % 17 == 0] = 2
    p = oe.params(**case)
    ghi, gfl, gst = oe.process(bases, offsets, paired, lo, hi, flags, p)
    ent = oe.values(bases, offsets, lo, hi, p.k, p.window)
    whi, wfl, wst = hi.copy(), flags.copy(), np.zeros(2, np.int64)
    cutoff = np.float32(max(0.0, p.cutoff))
    for u in range(0, len(lo), 2):
        if flags[u] & 2:
            continue
        d = []
        for i in (u, u + 1):
            disc = bool(flags[i] & 1)
            n = whi[i] - lo[i]
            if not (disc or (p.trim_failures_to_1bp and n == 1)):
                passes = bool(p.high_pass) ^ bool(ent[i] < cutoff)
                if not passes:
                    if p.trim_failures_to_1bp:
                        if n > 1:
                            whi[i] = lo[i] + 1
                    else:
                        disc = True
            wfl[i] = (wfl[i] & ~np.uint8(3)) | (1 if disc else 0)
            d.append(disc or (p.trim_failures_to_1bp and whi[i] - lo[i] == 1))
        remove = (p.remove_pairs_if_either_bad and any(d)) or all(d)
        if remove:
            wst += [2, (whi[u] - lo[u]) + (whi[u + 1] - lo[u + 1])]
            wfl[u] |= 2
            wfl[u + 1] |= 2
    assert np.array_equal(ghi, whi) and np.array_equal(gfl, wfl) and list(gst) == list(wst)
    if case.get("cutoff", 0) > 0 and case.get("high_pass", True):
        assert gst[0] > 20


def py_low_entropy_mask(seq, k, window, cutoff, high_pass=True):
    """positions maskLowEntropy marks (jgi/BBDuk.java:4432-4446): every full window without an undefined base whose entropy fails
    the cutoff. The window's entropy is computed FROM SCRATCH (counts of its k-mers, the table terms summed in sorted order), so
    only windows whose from-scratch and running-sum entropies could straddle the cutoff are ambiguous; those are reported."""
    n, wk = len(seq), window - k + 1
    marked, ambiguous = set(), set()
    if n < window:
        return marked, ambiguous
    codes = [CODE.get(b, 0) for b in seq]
    for i in range(window - 1, n):
        a = i - window + 1
        if any(b not in CODE for b in seq[a:i + 1]):
            continue
        c = {}
        for j in range(a + k - 1, i + 1):
            km = tuple(codes[j - k + 1:j + 1])
            c[km] = c.get(km, 0) + 1
        h = -sum((v / wk) * math.log(v / wk) for v in sorted(c.values())) / math.log(wk)
        h = max(h, 0.0)
        if abs(h - cutoff) < 1e-5:
            ambiguous.update(range(a, i + 1))
        if not (high_pass ^ (h < cutoff)):
            marked.update(range(a, i + 1))
    return marked, ambiguous


@pytest.mark.parametrize("case", [dict(cutoff=0.5), dict(cutoff=0.75, k=4, window=30), dict(cutoff=0.35, k=3, window=20, tf1=True),
                                  dict(cutoff=0.6, high_pass=False, k=2, window=12)])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_mask_and_trim_match_from_scratch_windows(case, mode):
    """entropymask / entropytrim: the oracle's BitSet against windows measured from scratch; counts and trims from the marks"""
    bases, offsets, lo, hi, flags = entropy_batch(300, 21, L=100)
    flags[0::2][np.arange(150) % 13 == 0] = 2
    flags[1::2][np.arange(150) % 13 == 0] = 2
    p = oe.params(**case)
    glo, ghi, bits, moff, st = oe.mask(bases, offsets, True, lo, hi, flags, p, mode)
    cutoff = float(np.float32(max(0.0, p.cutoff)))
    n_marked = reads = total = 0
    for i in range(len(lo)):
        seq = bytes(bases[offsets[i] + lo[i]:offsets[i] + hi[i]])
        n = len(seq)
        got = {j for j in range(n) if (int(bits[moff[i] + (j >> 5)]) >> (j & 31)) & 1}
        skip = bool(flags[i - i % 2] & 2) or bool(flags[i] & 1) or (p.trim_failures_to_1bp and n == 1)
        if skip:
            assert not got and glo[i] == lo[i] and ghi[i] == hi[i]
            continue
        want, amb = py_low_entropy_mask(seq, p.k, p.window, cutoff, bool(p.high_pass))
        if mode == 3:
            assert not got
            if amb:
                continue
            left = next((j for j in range(n) if j not in want), n)
            right = next((j for j in range(n) if (n - 1 - j) not in want), n)
            if left or right:
                if left + right + min(n, 1) > n:
                    left, right = 0, max(1, n - min(n, 1))
                assert (glo[i], ghi[i]) == (lo[i] + left, hi[i] - right), (i, seq)
                total += left + right
                reads += 1
            else:
                assert (glo[i], ghi[i]) == (lo[i], hi[i])
        else:
            assert got - amb == want - amb, (i, seq)
            n_marked += len(got)
            if not amb:
                x = sum(1 for j in got if seq[j:j + 1] != b"N" and (mode == 1 or not seq[j:j + 1].islower()))
                total += x
                reads += x > 0
    assert (n_marked > 500) if mode != 3 else (reads > 10)
    if mode != 3:
        assert st[1] >= total and st[0] >= reads  # (reads with an ambiguous window are not in our count)
    else:
        assert st[1] >= total and st[0] >= reads
