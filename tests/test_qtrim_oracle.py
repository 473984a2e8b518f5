"""CPU test of the poly-X / quality-trimming / quality-filter oracle (oracle/qtrim_oracle.c) against an independently
written Python restatement of jgi/BBDuk.java:2954-3052, :3074-3170, :4721-4825 + shared/TrimRead.java:140-169, :299-410
(poly-X runs found with regular expressions / string scans instead of the reference's counters), plus hand-checked cases."""
import math
import re
import numpy as np
import pytest

from oracle import qtrim as oq

F32 = np.float32
PE = np.array([10.0 ** (0 - .1 * i) for i in range(128)]).astype(F32)
PE[0], PE[1] = F32(.75), F32(.7)


def trim_e(q):
    if q <= 0:
        return F32(0.75)
    if q <= 1:
        return F32(0.75 - q * 0.05)
    return F32(min(0.7, 10.0 ** (-0.1 * q)))


def py_test_optimal(bases, q, e):
    """(left, right) of shared/TrimRead.java:348-410: Kadane over delta = e - probError, ties to the longer run"""
    n = len(bases)
    nprob = F32(max(min(F32(e * F32(1.1)), F32(1)), F32(0.75)))
    best, best_loc, best_cnt = F32(0), -1, -1
    score, count = F32(0), 0
    for i in range(n):
        pe = nprob if (bases[i] == ord("N") or q[i] < 1) else PE[q[i]]
        score = F32(score + F32(e - pe))
        if score > 0:
            count += 1
            if score > best or (score == best and count > best_cnt):
                best, best_cnt, best_loc = score, count, i
        else:
            score, count = F32(0), 0
    if best > 0:
        return best_loc - best_cnt + 1, n - best_loc - 1
    return 0, n


def trim_by_amount(lo, hi, left, right, m):
    left, right = max(left, 0), max(right, 0)
    n = hi - lo
    if n < 1:
        return lo, hi, 0
    m = min(n, max(m, 0))
    if left + right + m > n:
        right, left = max(1, n - m), 0
    return lo + left, hi - right, left + right


def py_run_ends(good, interval):
    """(bases to cut on the left, on the right) for testLeft / testRight and their N-only versions (shared/TrimRead.java:416-503):
    walk inwards from an end until `interval` good symbols in a row; cut through the last bad one met on the way"""
    n = len(good)

    def walk(seq):
        run, last_bad = 0, -1
        for i, g in enumerate(seq):
            if run >= interval:
                break
            if g:
                run += 1
            else:
                run, last_bad = 0, i
        return last_bad + 1
    return (walk(good), walk(good[::-1])) if n else (0, 0)


def py_trim_amounts(b, q, e, p):
    """the (left, right) amounts of TrimRead.trimFast for each of its three modes, with or without qualities"""
    n = len(b)
    tq = int(np.int8(int(p.trimq)))
    n_good = [c != ord("N") for c in b]
    if p.trim_mode == 0:
        if q is None:
            return (0, 0) if e >= 1 else py_run_ends(n_good, p.min_good_interval)
        return py_test_optimal(b, q, e)
    if p.trim_mode == 1:  # the first window of `window_length` qualities whose sum is below window * trimq: cut from its start
        w = p.window_length
        if q is None or n < w:
            return 0, (0 if tq > 0 else py_run_ends(n_good, p.min_good_interval)[1])
        thresh = max(w * tq, 1)
        for start in range(0, n - w + 1):
            if int(np.sum(q[start:start + w].astype(np.int64))) < thresh:
                return 0, n - start
        return 0, 0
    if q is None:
        return (0, 0) if tq < 0 else py_run_ends(n_good, p.min_good_interval)
    return py_run_ends([int(x) > tq for x in q], p.min_good_interval)


def py_detect_left(seq, min_poly, max_non, c):
    """jgi/BBDuk.java:4771-4791 on a bytes object: walk runs of c; a run of >= min_poly resets the error budget"""
    if len(seq) < min_poly:
        return 0
    trim_to, non, i = -1, 0, 0
    for m in re.finditer(b"[^" + c + b"]|" + c + b"+", seq):
        tok = m.group()
        if tok[:1] == c:
            if len(tok) >= min_poly:
                non, trim_to = 0, m.end() - 1
        else:
            non += 1
            if non > max_non:
                break
    return trim_to + 1


def py_detect_right(seq, min_poly, max_non, c):
    return py_detect_left(seq[::-1], min_poly, max_non, c)


def py_block(bases, quals, offsets, paired, lo, hi, flags, p):
    lo, hi, flags = lo.copy(), hi.copy(), flags.copy()
    st = np.zeros(8, np.int64)
    e = trim_e(p.trimq)
    per = 2 if paired else 1
    tf1, rieb = bool(p.trim_failures_to_1bp), bool(p.remove_pairs_if_either_bad)
    maxlen = p.max_read_length if p.max_read_length > 0 else 2 ** 31 - 1

    def disc(i, state):
        return state[i] or (tf1 and hi[i] - lo[i] == 1)

    for u in range(0, len(offsets) - per, per):
        if flags[u] & 2:
            continue
        idx = list(range(u, u + per))
        state = {i: bool(flags[i] & 1) for i in idx}
        qt = {i: False for i in idx}

        def set_discarded(i):
            if tf1:
                if hi[i] - lo[i] > 1:
                    lo[i], hi[i], _ = trim_by_amount(lo[i], hi[i], 0, hi[i] - lo[i] - 1, 1)
            else:
                state[i] = True

        def should_remove():
            d = [disc(i, state) for i in idx]
            return (rieb and any(d)) or all(d)

        pt = {i: False for i in idx}
        remove = False

        def seq_of(i):
            return bytes(bases[offsets[i] + lo[i]:offsets[i] + hi[i]])

        def minlen_of(i):
            L = offsets[i + 1] - offsets[i]
            return int(max(F32(F32(L) * F32(p.min_len_fraction)), F32(p.min_read_length)))

        def close_poly():
            if should_remove():
                st[7] += sum(hi[i] - lo[i] for i in idx)
                return True
            return False

        if p.trim_poly_a > 0:
            for i in idx:
                sq = seq_of(i)
                x = 0
                if len(sq) >= p.trim_poly_a:
                    left = max(len(sq) - len(sq.lstrip(b"A")), len(sq) - len(sq.lstrip(b"T")))
                    right = max(len(sq) - len(sq.rstrip(b"A")), len(sq) - len(sq.rstrip(b"T")))
                    left = left if left >= p.trim_poly_a else 0
                    right = right if right >= p.trim_poly_a else 0
                    if left or right:
                        lo[i], hi[i], x = trim_by_amount(lo[i], hi[i], left, right, 1)
                st[7] += x
                st[6] += x > 0
                pt[i] |= x > 0
                if hi[i] - lo[i] < minlen_of(i):
                    set_discarded(i)
            remove = close_poly()
        for c, tl, tr, fp in ((b"G", p.trim_poly_g_left, p.trim_poly_g_right, p.filter_poly_g),
                              (b"C", p.trim_poly_c_left, p.trim_poly_c_right, p.filter_poly_c)):
            if remove or not (tl > 0 or tr > 0 or fp > 0):
                continue
            for i in idx:
                probe = idx[0] if (c == b"C" and i != idx[0]) else i  # the reference tests r1 for r2's poly-C filter
                if fp > 0 and py_detect_left(seq_of(probe), fp, p.max_non_poly, c) >= fp:
                    set_discarded(i)
                    st[6] += 1
                elif tl > 0 or tr > 0:
                    sq = seq_of(i)
                    left = py_detect_left(sq, tl, p.max_non_poly, c) if tl > 0 else 0
                    right = py_detect_right(sq, tr, p.max_non_poly, c) if tr > 0 else 0
                    x = 0
                    if left or right:
                        lo[i], hi[i], x = trim_by_amount(lo[i], hi[i], left, right, 1)
                    st[7] += x
                    st[6] += x > 0
                    pt[i] |= x > 0
                    if hi[i] - lo[i] < minlen_of(i):
                        set_discarded(i)
            remove = close_poly()
        if remove:
            for i in idx:
                flags[i] = (flags[i] & ~np.uint8(3)) | (1 if state[i] else 0) | 2 | (0x80 if pt[i] else 0)
            continue
        if p.qtrim_left or p.qtrim_right:
            for i in idx:
                if hi[i] - lo[i] < 1:
                    continue
                b = bases[offsets[i] + lo[i]:offsets[i] + hi[i]]
                q = None if quals is None else (quals[offsets[i] + lo[i]:offsets[i] + hi[i]].astype(np.int64) - p.qual_offset).astype(np.int8)
                a0, b0 = py_trim_amounts(b, q, e, p)
                lo[i], hi[i], x = trim_by_amount(lo[i], hi[i], a0 if p.qtrim_left else 0, b0 if p.qtrim_right else 0, 1)
                st[1] += x
                st[0] += x > 0
                qt[i] = x > 0
        for i in idx:
            if not disc(i, state):
                n = hi[i] - lo[i]
                L = offsets[i + 1] - offsets[i]
                minlen = int(max(F32(F32(L) * F32(p.min_len_fraction)), F32(p.min_read_length)))
                if n < minlen or n > maxlen:
                    set_discarded(i)
        remove = False
        if should_remove():
            st[1] += sum(hi[i] - lo[i] for i in idx)
            remove = True
        if not remove:
            if p.min_avg_quality > 0:
                for i in idx:
                    n = hi[i] - lo[i]
                    if n == 0:
                        avg = 0.0
                    else:
                        lim = n if p.min_avg_quality_bases < 1 else min(n, p.min_avg_quality_bases)
                        b = bases[offsets[i] + lo[i]:offsets[i] + lo[i] + lim]
                        q = (quals[offsets[i] + lo[i]:offsets[i] + lo[i] + lim].astype(np.int64) - p.qual_offset).astype(np.int8)
                        ee = F32(0)
                        for bb, qq in zip(b, q):
                            if chr(bb) in "ACGTUacgtu":
                                ee = F32(ee + PE[max(int(qq), 0)])
                        pr = float(F32(ee / F32(lim)))
                        avg = 0.0 if pr >= 1 else 60.0 if pr <= 0.000001 else -10 * math.log10(pr)
                    if avg < float(F32(p.min_avg_quality)):
                        set_discarded(i)
            if p.min_base_quality > 0:
                for i in idx:
                    q = (quals[offsets[i] + lo[i]:offsets[i] + hi[i]].astype(np.int64) - p.qual_offset).astype(np.int8)
                    if min([41] + list(q)) < p.min_base_quality:
                        set_discarded(i)
            if p.max_ns >= 0:
                for i in idx:
                    b = bases[offsets[i] + lo[i]:offsets[i] + hi[i]]
                    n = sum(1 for c in b if chr(c) not in "ACGTUacgtu")
                    if n > p.max_ns:
                        st[4] += 1
                        st[5] += hi[i] - lo[i]
                        set_discarded(i)
            if float(F32(p.max_n_rate)) < 1:  # the raw discarded flag guards against a double count (jgi/BBDuk.java:3138-3149)
                for i in idx:
                    if state[i]:
                        continue
                    b = bases[offsets[i] + lo[i]:offsets[i] + hi[i]]
                    n = sum(1 for c in b if chr(c) not in "ACGTUacgtu")
                    if F32(n) > F32(F32(p.max_n_rate) * F32(len(b))):
                        st[4] += 1
                        st[5] += hi[i] - lo[i]
                        set_discarded(i)
            if p.min_consecutive_bases > 0:  # some run of that many defined bases (stream/Read.java:2846-2858)
                for i in idx:
                    if disc(i, state):
                        continue
                    b = bytes(bases[offsets[i] + lo[i]:offsets[i] + hi[i]]).decode("latin1")
                    runs = "".join("x" if c in "ACGTUacgtu" else " " for c in b).split()
                    if max([0] + [len(r) for r in runs]) < p.min_consecutive_bases:
                        set_discarded(i)
            if float(F32(p.min_base_frequency)) > 0:  # the rarest of upper-case A C G T (stream/Read.java:2864-2874)
                for i in idx:
                    b = bytes(bases[offsets[i] + lo[i]:offsets[i] + hi[i]])
                    mn = min(b.count(c) for c in (b"A", b"C", b"G", b"T"))
                    if F32(mn) < F32(F32(p.min_base_frequency) * F32(len(b))):
                        set_discarded(i)
            if should_remove():
                st[3] += sum(hi[i] - lo[i] for i in idx)
                st[2] += per
                remove = True
        for i in idx:
            flags[i] = (flags[i] & ~np.uint8(3)) | (1 if state[i] else 0) | (2 if remove else 0) | (0x40 if qt[i] else 0) | \
                (0x80 if pt[i] else 0)
    return lo, hi, flags, st


def qual_batch(n, seed, L=80, paired=True):
    """reads with quality profiles that decay, recover, dip in the middle; N's; a fake k-mer block in front"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    lens = rng.integers(0, L + 1, n)
    lens[rng.random(n) < 0.5] = L
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    bases = acgt[rng.integers(0, 4, int(offsets[-1]))].copy()
    quals = np.zeros(len(bases), np.int64)
    for i in range(n):
        ln = int(lens[i])
        kind = rng.integers(0, 6)
        x = np.arange(ln)
        if kind == 0:
            q = rng.integers(25, 42, ln)
        elif kind == 1:
            q = np.clip(40 - (x * rng.integers(20, 60)) // max(ln, 1) + rng.integers(-4, 5, ln), 0, 41)
        elif kind == 2:
            q = np.clip(5 + (x * rng.integers(20, 50)) // max(ln, 1) + rng.integers(-4, 5, ln), 0, 41)
        elif kind == 3:
            q = rng.integers(30, 42, ln)
            if ln > 10:
                a = int(rng.integers(0, ln - 5))
                q[a:a + int(rng.integers(1, 12))] = rng.integers(0, 8)
        elif kind == 4:
            q = rng.integers(0, 12, ln)
        else:
            q = rng.integers(0, 42, ln)
        quals[offsets[i]:offsets[i + 1]] = q
    for i in range(n):  # homopolymer heads / tails, some with an interruption
        ln = int(lens[i])
        if ln < 12:
            continue
        for side in (0, 1):
            if rng.random() < 0.25:
                k = int(rng.integers(1, 14))
                c = rng.choice(np.frombuffer(b"AATGGC", np.uint8))
                seg = slice(offsets[i], offsets[i] + k) if side == 0 else slice(offsets[i + 1] - k, offsets[i + 1])
                bases[seg] = c
                if k > 4 and rng.random() < 0.4:
                    bases[seg.start + int(rng.integers(1, k - 1))] = ord("A") if c != ord("A") else ord("C")
    nn = rng.random(len(bases)) < 0.02
    bases[nn] = ord("N")
    odd = rng.random(len(bases)) < 0.003
    bases[odd] = rng.choice(np.frombuffer(b"acgtnRY", np.uint8), int(odd.sum()))
    quals[nn & (rng.random(len(bases)) < 0.5)] = 0
    lo = np.zeros(n, np.int32)
    hi = lens.astype(np.int32)
    cut = rng.random(n) < 0.3
    hi[cut] = np.maximum(0, hi[cut] - rng.integers(1, 30, int(cut.sum()))).astype(np.int32)
    cutl = rng.random(n) < 0.1
    lo[cutl] = np.minimum(hi[cutl], rng.integers(1, 10, int(cutl.sum()))).astype(np.int32)
    flags = np.zeros(n, np.uint8)
    per = 2 if paired else 1
    rem = rng.random(n // per) < 0.05
    for q in range(per):
        flags[q::per][rem[:len(flags[q::per])]] = 2
    d = rng.random(n) < 0.03
    flags[d & (flags == 0)] = 1
    return bases, (quals + 33).astype(np.uint8), offsets, lo, hi, flags


CASES = [dict(qtrim="rl", trimq=10.0), dict(qtrim="r", trimq=6.0), dict(qtrim="l", trimq=15.5, minlen=25),
         dict(qtrim="rl", trimq=20.0, rieb=False), dict(qtrim="rl", trimq=10.0, tf1=True, mbq=3),
         dict(qtrim="", mbq=5, maxns=1), dict(qtrim="r", trimq=0.5, maxns=0, maxlen=70, mlf=0.5), dict(qtrim="rl", trimq=1.0),
         dict(qtrim="", polya=3, minlen=30), dict(qtrim="rl", trimq=8.0, polyg=(4, 4), fpolyc=5, maxnonpoly=1),
         dict(qtrim="", polyg=(2, 0), polyc=(0, 3), fpolyg=6, maxnonpoly=0, rieb=False),
         dict(qtrim="r", trimq=12.0, polya=2, polyg=(3, 3), polyc=(3, 3), fpolyg=8, fpolyc=8, maxnonpoly=2, tf1=True),
         dict(qtrim="", maq=20.0), dict(qtrim="rl", trimq=5.0, maq=13.5, maqb=40, mbq=1), dict(qtrim="", maq=7.0, rieb=False),
         dict(qtrim="", maxnrate=0.02), dict(qtrim="r", trimq=8.0, maxns=1, maxnrate=0.01, rieb=False), dict(qtrim="", mcb=30),
         dict(qtrim="rl", trimq=10.0, mcb=12, tf1=True), dict(qtrim="", mbf=0.18), dict(qtrim="r", trimq=6.0, mbf=0.1, mcb=20, maxnrate=0.05),
         dict(qtrim="r", trimq=12.0, mode=1), dict(qtrim="r", trimq=20.0, mode=1, window=7, rieb=False), dict(qtrim="rl", trimq=10.0, mode=2),
         dict(qtrim="l", trimq=15.0, mode=2, goodinterval=4, minlen=20), dict(qtrim="r", trimq=8.9, mode=2, goodinterval=1)]
# the same rules on reads WITHOUT qualities (N's only), run with quals=None
NOQUAL_CASES = [dict(qtrim="rl", trimq=10.0), dict(qtrim="r", trimq=6.0, mode=2), dict(qtrim="rl", trimq=0.0, mode=1), dict(qtrim="l", trimq=5.0, mode=2, goodinterval=3),
                dict(qtrim="rl", trimq=-1.0, mode=2), dict(qtrim="rl", trimq=10.0, maxns=2, mcb=25)]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_oracle_matches_python_restatement(case):
    paired = case % 2 == 0
    bases, quals, offsets, lo, hi, flags = qual_batch(400, 100 + case, paired=paired)
    p = oq.params(**CASES[case])
    got = oq.process(bases, quals, offsets, paired, lo, hi, flags, p)
    want = py_block(bases, quals, offsets, paired, lo, hi, flags, p)
    for g, w, name in zip(got, want, ("lo", "hi", "flags", "stats")):
        assert np.array_equal(g, w), name
    assert got[3].sum() > 0


@pytest.mark.parametrize("case", range(len(NOQUAL_CASES)))
def test_oracle_matches_python_restatement_without_qualities(case):
    """reads without qualities: every trimming rule falls back to trimming N's (shared/TrimRead.java:352, :418, :440, :459)"""
    paired = case % 2 == 1
    bases, _, offsets, lo, hi, flags = qual_batch(400, 300 + case, paired=paired)
    rng = np.random.default_rng(case)
    ends = rng.random(len(offsets) - 1) < 0.5  # N's at the read ends, where the N-only rules look
    for i in np.nonzero(ends)[0]:
        a, b = offsets[i], offsets[i + 1]
        k = int(rng.integers(0, 6))
        bases[a:a + min(k, b - a)] = ord("N")
        bases[max(a, b - int(rng.integers(0, 6))):b] = ord("N")
    p = oq.params(**NOQUAL_CASES[case])
    got = oq.process(bases, None, offsets, paired, lo, hi, flags, p)
    want = py_block(bases, None, offsets, paired, lo, hi, flags, p)
    for g, w, name in zip(got, want, ("lo", "hi", "flags", "stats")):
        assert np.array_equal(g, w), name


def test_trim_mode_known_answers():
    """hand-checked cases of the window rule and of testLeft / testRight"""
    def one(qs, **kw):
        n = len(qs)
        b = np.frombuffer(("A" * n).encode(), np.uint8)
        q = (np.array(qs) + 33).astype(np.uint8)
        off = np.array([0, n], np.int64)
        lo, hi, _, _ = oq.process(b, q, off, False, np.zeros(1, np.int32), np.array([n], np.int32), np.zeros(1, np.uint8),
                                  oq.params(minlen=1, **kw))
        return int(lo[0]), int(hi[0])
    # window of 4 below 4 * 10 first at positions 12..15 (30 30 2 2 = 64 >= 40; 30 2 2 2 = 36 < 40 starts at 11)
    assert one([30] * 12 + [2] * 8, qtrim="r", trimq=10.0, mode=1) == (0, 11)
    assert one([30] * 20, qtrim="r", trimq=10.0, mode=1) == (0, 20)
    assert one([5, 5, 5, 5] + [30] * 10, qtrim="r", trimq=10.0, mode=1) == (0, 1)  # the very first window fails: one base is kept
    # optitrim=f: walk in from each end until 2 good bases in a row; cut through the last bad one seen
    assert one([2, 30, 2, 30, 30, 30, 30, 2, 30, 30], qtrim="rl", trimq=10.0, mode=2) == (3, 10)
    assert one([30, 30, 30, 2, 30], qtrim="rl", trimq=10.0, mode=2) == (0, 3)
    assert one([30, 2, 30, 30], qtrim="l", trimq=10.0, mode=2, goodinterval=1) == (0, 4)


def test_known_answers():
    """hand-checked cases of testOptimal + trimByAmount"""
    def one(qs, seq=None, **kw):
        n = len(qs)
        b = np.frombuffer((seq or "A" * n).encode(), np.uint8)
        q = (np.array(qs) + 33).astype(np.uint8)
        off = np.array([0, n], np.int64)
        lo, hi, fl, st = oq.process(b, q, off, False, np.zeros(1, np.int32), np.array([n], np.int32), np.zeros(1, np.uint8),
                                    oq.params(minlen=1, **kw))
        return int(lo[0]), int(hi[0]), int(fl[0]), list(st)
    # all bases better than trimq: nothing trimmed
    assert one([30] * 20, trimq=10.0) == (0, 20, 0, [0, 0, 0, 0, 0, 0, 0, 0])
    # a bad tail: the five Q2 bases go
    assert one([30] * 15 + [2] * 5, trimq=10.0) == (0, 15, 0x40, [1, 5, 0, 0, 0, 0, 0, 0])
    # bad head and tail, qtrim=r only trims the tail
    assert one([2] * 4 + [30] * 10 + [2] * 6, qtrim="r", trimq=10.0) == (0, 14, 0x40, [1, 6, 0, 0, 0, 0, 0, 0])
    assert one([2] * 4 + [30] * 10 + [2] * 6, qtrim="rl", trimq=10.0) == (4, 14, 0x40, [1, 10, 0, 0, 0, 0, 0, 0])
    # everything bad: trimByAmount keeps one base (right = len - 1), then minlen=1 keeps the read
    assert one([2] * 12, trimq=10.0) == (0, 1, 0x40, [1, 11, 0, 0, 0, 0, 0, 0])
    # an N costs 0.75 - 0.1 = 6.5 good bases: after 5 good bases the run dies and the longer side is kept, after 8 it survives
    lo, hi, fl, st = one([35] * 30, seq="A" * 5 + "N" + "A" * 24, trimq=10.0)
    assert (lo, hi) == (6, 30)
    lo, hi, fl, st = one([35] * 30, seq="A" * 8 + "N" + "A" * 21, trimq=10.0)
    assert (lo, hi) == (0, 30)
    # poly-X: trimpolya takes A or T runs of >= 3 from both ends; trimpolyg tolerates one non-G inside the tail
    assert one([35] * 20, seq="TTTTCGCGCGCGCGCGCAAA", qtrim="", polya=3) == (4, 17, 0x80, [0, 0, 0, 0, 0, 0, 1, 7])
    assert one([35] * 20, seq="TTCGCGCGCGCGCGCGCGAA", qtrim="", polya=3) == (0, 20, 0, [0] * 8)
    assert one([35] * 20, seq="ACGTACGTACGTGGGGAGGG", qtrim="", polyg=(0, 3), maxnonpoly=1) == (0, 12, 0x80, [0, 0, 0, 0, 0, 0, 1, 8])
    assert one([35] * 20, seq="ACGTACGTACGTGGGGAGGG", qtrim="", polyg=(0, 3), maxnonpoly=0) == (0, 17, 0x80, [0, 0, 0, 0, 0, 0, 1, 3])
    # filterpolyg discards (flag 1, unit removed 2) and counts the read, not its bases... the pair length goes to basesPoly
    assert one([35] * 20, seq="GGGGGGGGACGTACGTACGT", qtrim="", fpolyg=6) == (0, 20, 0x03, [0, 0, 0, 0, 0, 0, 1, 20])
