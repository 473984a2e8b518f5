"""TEST INFRASTRUCTURE: second, independently written restatement of the BBDuk k-mer path.

Where oracle/bbduk_oracle.c follows the reference's serial loops statement by statement, this file
states the same semantics position by position in closed form with
strings and Python ints, and builds the reference table as an explicit Hamming ball instead of the
mutate() recursion. The two must agree on every input (tests/test_oracle.py); that agreement, the
Appendix-B key counts and the reference's own asserts are what pins the oracle while no JVM is
available (PARITY UNPINNED by reference-run vectors).

Pure Python, small inputs only. Substitution neighbourhoods only (edist=0). The first half (hit_id, ktrim_right,
ktrim_left, kfilter_count, kmask_bits) takes qhdist=0, speed=0, qskip=1; the second half (probe and everything after
it) adds query-side substitutions, qskip, speed= of both generations, and the modes the tuned GPU kernels do not serve
(ktrimTips, ksplit, countCoveredBases, findBestMatch, countSetKmersBig), so that probe_generic.cu is not checked
only against its twin transliteration in bbduk_oracle.c.
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

    def __init__(self, k=27, mink=-1, hdist=0, hdist2=None, mm=True, rcomp=True, fn=False, generation=0,
                 qhdist=0, qhdist2=None, qskip=1, speed=0, trimming=True):
        self.kbig = k if (k > 31 and not trimming and speed < 1 and qskip < 2) else 0  # jgi/BBDuk.java:709, :765-779
        if self.kbig:
            mm = False  # :796-800
        self.qhdist = qhdist
        self.qhdist2 = qhdist if qhdist2 is None else qhdist2
        self.qskip = qskip
        self.speed = speed
        self.generation = generation
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
            if dist == 0 and not passes_speed(d, kk):
                continue  # the loader applies speed= only where it stores without mutating (jgi/BBDuk.java:2365 vs :2392-2401)
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


# ---------------------------------------------------------------------------------------------------------------------
# Second half: query-side options and the modes served only by the generic GPU kernel.
# ---------------------------------------------------------------------------------------------------------------------
M64 = (1 << 64) - 1


def passes_speed(d, key):
    """jgi.BBDuk: (key & Long.MAX_VALUE) % 17 >= speed (jgi/BBDuk.java:4702-4713); bbduk.BBDukS with its default index:
    ((hash64plus2(key) >> 16) & 15) + 1 >= speed (bbduk/BBDukIndexMask2.java:566-577, shared/Tools.java:5482-5497)"""
    if d.generation == 0:
        return d.speed < 1 or (key & ((1 << 63) - 1)) % 17 >= d.speed
    if d.speed < 2:
        return True
    h = key & M64
    h ^= h >> 33
    h = (h * 0xff51afd7ed558ccd) & M64
    h ^= h >> 33
    h = (h * 0xc4ceb9fe1a85ec53) & M64
    h ^= h >> 33
    h &= (1 << 63) - 1
    if h >= 0x7FFFF800FFFFFFFF:
        h = ((h - 0x7FFFF800FFFFFFFF) * 64) & M64
        if h >> 63:
            h -= 1 << 64  # Java long
    return ((h >> 16) & 15) + 1 >= d.speed


def _mask_mid(d, v, n):
    if d.mml and n == d.k:
        for s_ in range(d.mm_lo, d.mm_lo + d.mml):
            v &= ~(3 << (2 * s_))
    return v


def _lookup(d, table, kmer, rkmer, n, qpos):
    """getValueInner (jgi/BBDuk.java:3365-3386)"""
    if d.qskip > 1 and qpos % d.qskip != 0:
        return -1
    v = max(kmer, rkmer) if d.rcomp else kmer
    key = _mask_mid(d, v, n) | (1 << (2 * n))
    if not passes_speed(d, key):
        return -1
    return table.get(key, -1)


def _rc_int(kmer, n):
    out = 0
    for _ in range(n):
        out = (out << 2) | (3 - (kmer & 3))
        kmer >>= 2
    return out


def _get_value(d, table, kmer, rkmer, n, qpos, dist):
    """getValue (jgi/BBDuk.java:3335-3354): the exact k-mer, else the first hit among its substitution variants in the
    order symbol-major, slot-minor (slot 0 = last base), depth first; a variant's reverse k-mer is recomputed from it."""
    hit = _lookup(d, table, kmer, rkmer, n, qpos)
    if hit >= 1 or dist < 1:
        return hit
    for sym in range(4):
        for slot in range(n):
            t = (kmer & ~(3 << (2 * slot))) | (sym << (2 * slot))
            if t != kmer:
                hit = _get_value(d, table, t, _rc_int(t, n), n, qpos, dist - 1)
                if hit >= 1:
                    return hit
    return hit


def probe(d, table, seq, i, start=0):
    """id of the full-length probe at read position i of a scan that began at `start` with an empty state, -1 for a miss,
    None when the reference makes no probe there. The window reaches back to i-k+1 even in front of `start`: the bases the
    scan has not seen are zero bits in both registers (A.2)."""
    k = d.k
    if i < k - 1 or i < start:
        return None
    p = start - 1
    if d.forbidNs:
        for j in range(i, start - 1, -1):
            if not defined(seq[j]):
                p = j
                break
    if i - p < d.minlen2:
        return None
    lo = max(start, i - k + 1)
    kmer = pack([code0(seq[j]) for j in range(lo, i + 1)])
    rk = 0
    for j in range(max(p + 1, lo), i + 1):
        rk |= comp0(seq[j]) << (2 * (k - 1 - (i - j)))
    return _get_value(d, table, kmer, rk, k, i, d.qhdist)


def tail_probe(d, table, sub, qpos):
    """a short k-mer (the tails read undefined bases as code 0 on both strands)"""
    f = pack([code0(c) for c in sub])
    r = pack([comp0(c) for c in reversed(sub)])
    return _get_value(d, table, f, r, len(sub), qpos, d.qhdist2)


def _guard(d, L):
    return L >= max(1, min(d.k, d.mink) if d.usk else d.k)


def ktrim_tip(d, table, seq, start, stop, right):
    """ktrimTip (jgi/BBDuk.java:3706-3858), trimpad 0, inclusive -> (kept lo, kept hi, id0 or -1) on the string given"""
    L = len(seq)
    if not _guard(d, L) or not table:
        return 0, L, -1
    hits = [(i, probe(d, table, seq, i, start)) for i in range(start, stop)]
    hits = [(i, h) for i, h in hits if h is not None and h > 0]
    if hits:
        id0 = hits[0][1]
        min_loc = min(i - d.k + 1 for i, _ in hits)
        max_loc = hits[-1][0]
    elif d.usk:
        found = []
        if right:  # suffixes of seq[:stop], growing
            for n in range(1, min(d.k - 1, stop) + 1):
                if n >= d.mink:
                    h = tail_probe(d, table, seq[stop - n:stop], stop - n)
                    if h > 0:
                        found.append((stop - n, h))
            if not found:
                return 0, L, -1
            id0, min_loc, max_loc = found[0][1], found[-1][0], L - 1
        else:  # prefixes of seq[start:], growing, up to position min(k, stop)-1
            for i in range(start, min(d.k, stop)):
                n = i - start + 1
                if n >= d.mink:
                    h = tail_probe(d, table, seq[start:i + 1], i)
                    if h > 0:
                        found.append((i, h))
            if not found:
                return 0, L, -1
            id0, min_loc, max_loc = found[0][1], 0, max(i for i, _ in found)
    else:
        return 0, L, -1
    if right:
        lo, hi = trim_by_amount(L, 0, L - (min_loc - 1) - 1)
    else:
        lo, hi = trim_by_amount(L, max_loc + 1, 0)
    return lo, hi, id0


def ktrim_tips(d, table, seq, restrict_left=0, restrict_right=0):
    """ktrimTips (jgi/BBDuk.java:3686-3699): the right half first, then the left half OF WHAT IS LEFT, both split at the
    original middle -> (kept lo, kept hi, id credited by the right pass, id credited by the left pass, bases trimmed)"""
    L = len(seq)
    mid = L // 2 - (d.k - 1) // 2
    start = max(0, mid if restrict_right < 1 else L - restrict_right)
    _, hi, id_r = ktrim_tip(d, table, seq, start, L, True)
    seq2 = seq[:hi]
    stop = min(len(seq2), mid + d.k - 1 if restrict_left < 1 else restrict_left)
    lo, hi2, id_l = ktrim_tip(d, table, seq2, 0, stop, False)
    return lo, hi2, id_r, id_l, (L - hi) + (len(seq2) - (hi2 - lo))


def ksplit(d, table, seq):
    """ksplit (jgi/BBDuk.java:4208-4377), trimpad 0 -> (kept lo, kept hi, split?, start of the new mate, id0)"""
    L = len(seq)
    if not _guard(d, L) or not table or L < d.k:
        return 0, L, False, -1, -1
    hits = [(i, probe(d, table, seq, i)) for i in range(L)]
    hits = [(i, h) for i, h in hits if h is not None and h > 0]
    if hits:
        id0 = hits[0][1]
        leftmost = min(max(0, i - (d.k - 1)) for i, _ in hits)
        rightmost = max(i for i, _ in hits)
    elif d.usk:
        right = []
        for n in range(1, min(d.k - 1, L) + 1):
            if n >= d.mink:
                h = tail_probe(d, table, seq[L - n:], L - n)
                if h > 0:
                    right.append((L - n, h))
        if right:
            id0, leftmost, rightmost = right[0][1], min(i for i, _ in right), L - 1
        else:
            left = []
            for n in range(1, min(d.k, L) + 1):
                if n >= d.mink:
                    h = tail_probe(d, table, seq[:n], n - 1)
                    if h > 0:
                        left.append((n - 1, h))
            if not left:
                return 0, L, False, -1, -1
            id0, leftmost, rightmost = left[0][1], 0, max(i for i, _ in left)
    else:
        return 0, L, False, -1, -1
    if leftmost == 0:
        lo, hi = trim_by_amount(L, rightmost + 1, 0)
        return lo, hi, False, -1, id0
    if rightmost == L - 1:
        lo, hi = trim_by_amount(L, 0, L - (leftmost - 1) - 1)
        return lo, hi, False, -1, id0
    lo, hi = trim_by_amount(L, 0, L - (leftmost - 1) - 1)
    return lo, hi, True, rightmost + 1, id0


def covered_bases(d, table, seq, min_covered):
    """countCoveredBases (jgi/BBDuk.java:3466-3519), no hit-count histogram -> (returned count, credited id or -1)"""
    L = len(seq)
    if L < d.k or not table:
        return 0, -1
    found, last = 0, -1
    for i in range(L):
        h = probe(d, table, seq, i)
        if h is not None and h > 0:
            found += min(d.k, i - last)
            last = i
            if found >= min_covered:
                return found, h
    return found, -1


def best_match(d, table, seq, max_bad=0):
    """findBestMatch (jgi/BBDuk.java:3527-3589) -> id with the most hits (ties: the id seen first) if hits > max_bad else -1"""
    L = len(seq)
    if L < d.k or not table:
        return -1
    order, counts = [], {}
    for i in range(L):
        h = probe(d, table, seq, i)
        if h is not None and h > 0:
            if h not in counts:
                order.append(h)
                counts[h] = 0
            counts[h] += 1
    if sum(counts.values()) <= max_bad:
        return -1
    top = max(counts.values())
    return next(h for h in order if counts[h] == top)


def count_big(d, table, seq, max_bad=0):
    """countSetKmersBig (jgi/BBDuk.java:3596-3677): k > 31 as runs of consecutive 31-mer hits. A run is a maximal stretch of
    PROBES that hit (positions without a probe neither extend nor end it); a run from position a to b adds b-a-(kbig-k-1)
    when that is positive; early return once the sum exceeds max_bad -> (returned count, credited id or -1)"""
    L = len(seq)
    if L < d.kbig or not table:
        return 0, -1
    sub = d.kbig - d.k - 1
    found, cred = 0, -1
    run = None  # (first position, last position, last id)
    last_id = -1

    def close(found, run):
        dif = run[1] - run[0] - sub
        return found + dif if dif > 0 else found

    for i in range(L):
        h = probe(d, table, seq, i)
        if h is None:
            continue
        if h > 0:
            last_id = h
            run = (run[0] if run else i, i)
        elif run:
            old = found
            found = close(found, run)
            run = None
            if found > max_bad >= old:
                return found, last_id
    if run:
        old = found
        found = close(found, run)
        if found > max_bad >= old:
            cred = last_id
    return found, cred
