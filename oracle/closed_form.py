"""TEST INFRASTRUCTURE: second, independently written restatement of the BBDuk k-mer path.

Where oracle/bbduk_oracle.c follows the reference's serial loops statement by statement, this file
states the same semantics position by position in closed form with
strings and Python ints, and builds the reference table as an explicit Hamming ball instead of the
mutate() recursion. The two must agree on every input (tests/test_oracle.py); that agreement, the
Appendix-B key counts and the reference's own asserts are what pins the oracle while no JVM is
available (PARITY UNPINNED by reference-run vectors).

Pure Python, small inputs only. Substitution neighbourhoods only (edist=0), qhdist=0, speed=0, qskip=1.
"""
from itertools import combinations, product

CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}


def defined(ch):
    return ch.upper() in CODE


def code0(ch):
    return CODE.get(ch.upper(), 0)


def comp0(ch):
    return 3 - CODE[ch.upper()] if defined(ch) else 0


def pack(codes):
    v = 0
    for c in codes:
        v = (v << 2) | c
    return v


def rc_codes(codes):
    return [3 - c for c in reversed(codes)]


class Derived:
    """jgi/BBDuk.java:672-877, only what the closed form needs."""

    def __init__(self, k=27, mink=-1, hdist=0, hdist2=None, mm=True, rcomp=True, fn=False, generation=0):
        self.k = min(k, 31)
        self.hdist = hdist
        self.hdist2 = hdist if hdist2 is None else hdist2
        self.rcomp = rcomp
        self.forbidNs = fn or hdist < 1
        mml = (2 - (self.k & 1)) if mm else 0
        self.minlen2 = (self.k - mml) // 2 if mm else self.k
        self.usk = 0 < mink < self.k
        self.mink = min(6 if (mink < 1 and generation == 0) else mink, self.k)
        if self.usk:
            mml = 0
        self.mml = mml
        self.mm_lo = (self.k - mml) // 2 if mml else None  # masked slots [mm_lo, mm_lo+mml), slot 0 = last base


def key_of(d, codes):
    """canonical key of a fully specified k-mer (list of codes, first base first): (max(fwd, rc) & middleMask) | 1<<2len"""
    n = len(codes)
    f, r = pack(codes), pack(rc_codes(codes))
    v = max(f, r) if d.rcomp else f
    if d.mml and n == d.k:
        for s in range(d.mm_lo, d.mm_lo + d.mml):
            v &= ~(3 << (2 * s))
    return v | (1 << (2 * n))


def ball(codes, dist):
    """all code lists within Hamming distance <= dist"""
    n = len(codes)
    for r in range(dist + 1):
        for pos in combinations(range(n), r):
            for subs in product(range(1, 4), repeat=r):
                c = list(codes)
                for p_, s in zip(pos, subs):
                    c[p_] = (c[p_] + s) & 3
                yield c


def build_table(d, scaffolds):
    """key -> smallest 1-based scaffold id (jgi/BBDuk.java:2210-2452 as set semantics)"""
    table = {}

    def put(codes, dist, sid):
        for c in ball(codes, dist):
            kk = key_of(d, c)
            if kk not in table or table[kk] > sid:
                table[kk] = sid

    for sid, seq in enumerate(scaffolds, 1):
        L = len(seq)
        if L < d.k:
            continue
        for i in range(d.k - 1, L):
            win = seq[i - d.k + 1:i + 1]
            if not all(defined(ch) for ch in win):
                continue
            codes = [code0(ch) for ch in win]
            put(codes, d.hdist, sid)
            if d.usk:
                if i == d.k - 1:
                    for n in range(d.k - 1, d.mink - 1, -1):
                        put(codes[:n], d.hdist2, sid)
                if i == L - 1:
                    for n in range(d.k - 1, d.mink - 1, -1):
                        put(codes[d.k - n:], d.hdist2, sid)
    return table


def hit_id(d, table, seq, i, start=0):
    """id (>0) of the full-length probe at read position i, or -1; None if no probe happens there (A.2/A.3)."""
    k = d.k
    # p = last undefined position <= i that resets (forbidNs), else start-1
    p = start - 1
    if d.forbidNs:
        for j in range(i, start - 1, -1):
            if not defined(seq[j]):
                p = j
                break
    ln = i - p
    if not (ln >= d.minlen2 and i >= k - 1):
        return None
    lo = max(start, i - k + 1)
    kmer = pack([code0(seq[j]) for j in range(lo, i + 1)])
    rk = 0
    for j in range(max(p + 1, lo), i + 1):
        rk |= comp0(seq[j]) << (2 * (k - 1 - (i - j)))
    v = max(kmer, rk) if d.rcomp else kmer
    if d.mml:
        for s in range(d.mm_lo, d.mm_lo + d.mml):
            v &= ~(3 << (2 * s))
    return table.get(v | (1 << (2 * k)), -1)


def tail_id(d, table, codes_fwd, codes_rc):
    n = len(codes_fwd)
    f, r = pack(codes_fwd), pack(codes_rc)
    v = max(f, r) if d.rcomp else f
    return table.get(v | (1 << (2 * n)), -1)


def trim_by_amount(L, left, right, min_len=1):
    """shared/TrimRead.java:299-346 -> kept [lo, hi)"""
    left, right = max(left, 0), max(right, 0)
    if L < 1:
        return 0, L
    m = min(L, max(min_len, 0))
    if left + right + m > L:
        right, left = max(1, L - m), 0
    return left, L - right


def ktrim_right(d, table, seq):
    """A.5 for ktrim=r, trimpad 0, inclusive: -> (kept_hi, id0) ; id0=-1 when nothing found"""
    L = len(seq)
    if L < max(1, min(d.k, d.mink) if d.usk else d.k) or not table:
        return L, -1
    hits = [(i, hit_id(d, table, seq, i)) for i in range(L)]
    hits = [(i, h) for i, h in hits if h is not None and h > 0]
    if hits:
        min_loc = min(i - d.k + 1 for i, _ in hits)
        id0 = hits[0][1]
    elif d.usk:
        found = []
        for n in range(1, min(d.k - 1, L) + 1):  # suffix of length n, scan order = growing n
            if n >= d.mink:
                sub = seq[L - n:]
                f = [code0(ch) for ch in sub]
                r = [comp0(ch) for ch in reversed(sub)]
                h = tail_id(d, table, f, r)
                if h > 0:
                    found.append((L - n, h))
        if not found:
            return L, -1
        id0 = found[0][1]
        min_loc = found[-1][0]
    else:
        return L, -1
    lo, hi = trim_by_amount(L, 0, L - (min_loc - 1) - 1)
    return hi, id0


def ktrim_left(d, table, seq):
    """A.5 for ktrim=l -> (kept_lo, kept_hi, id0)"""
    L = len(seq)
    if L < max(1, min(d.k, d.mink) if d.usk else d.k) or not table:
        return 0, L, -1
    hits = [(i, hit_id(d, table, seq, i)) for i in range(L)]
    hits = [(i, h) for i, h in hits if h is not None and h > 0]
    if hits:
        max_loc = hits[-1][0]
        id0 = hits[0][1]
    elif d.usk:
        found = []
        for n in range(1, min(d.k, L) + 1):  # prefix of length n (n may reach k, jgi/BBDuk.java:3914)
            if n >= d.mink:
                sub = seq[:n]
                f = [code0(ch) for ch in sub]
                r = [comp0(ch) for ch in reversed(sub)]
                h = tail_id(d, table, f, r)
                if h > 0:
                    found.append((n - 1, h))
        if not found:
            return 0, L, -1
        id0 = found[0][1]
        max_loc = max(i for i, _ in found)
    else:
        return 0, L, -1
    lo, hi = trim_by_amount(L, max_loc + 1, L - (L - 1) - 1)
    return lo, hi, id0


def kfilter_count(d, table, seq, max_bad=0):
    """A.7 countSetKmers -> (returned count, credited id or -1)"""
    L = len(seq)
    if L < d.k or not table:
        return 0, -1
    found = 0
    for i in range(L):
        h = hit_id(d, table, seq, i)
        if h is not None and h > 0:
            if found == max_bad:
                return found + 1, h
            found += 1
    return found, -1


def kmask_bits(d, table, seq):
    """A.7c kmask, trimpad 0, mfc off -> (set of masked positions, id0)"""
    L = len(seq)
    if L < d.k or not table:
        return set(), -1
    bits, ids = set(), []
    for i in range(L):
        h = hit_id(d, table, seq, i)
        if i >= d.k - 1 and h is not None and h > 0:
            ids.append(h)
            bits.update(range(max(0, i - (d.k - 1)), i + 1))
    if d.usk:
        for n in range(1, min(d.k, L) + 1):
            if n >= d.mink:
                sub = seq[:n]
                h = tail_id(d, table, [code0(c) for c in sub], [comp0(c) for c in reversed(sub)])
                if h > 0:
                    ids.append(h)
                    bits.update(range(0, min(L, n)))
        for n in range(1, min(d.k - 1, L) + 1):
            if n >= d.mink:
                sub = seq[L - n:]
                h = tail_id(d, table, [code0(c) for c in sub], [comp0(c) for c in reversed(sub)])
                if h > 0:
                    ids.append(h)
                    bits.update(range(L - n, L))
    if not ids:
        return set(), -1
    return bits, ids[0]
