"""ctypes binding of the Seal oracle (oracle/seal_oracle.c inside libbbduk_oracle.so) and an independent
closed-form Python restatement of the same path (no rolling state: every k-mer is read off the string).

TEST INFRASTRUCTURE ONLY -- never imported by bbtools_b200. PARITY UNPINNED (see seal_oracle.c)."""
import ctypes as C
import itertools
import math

import numpy as np

from .oracle import build

from bbtools_b200.seal import (AMBIG_ALL, AMBIG_FIRST, AMBIG_RANDOM, AMBIG_TOSS, MATCH_ALL, MATCH_FIRST, MATCH_UNIQUE, SealCfg, SealOut,  # noqa: F401
                               SealResult, SealStats, make_cfg, n_units)

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.sl_ora_create.restype = C.c_void_p
        L.sl_ora_create.argtypes = [C.POINTER(SealCfg)]
        L.sl_ora_destroy.argtypes = [C.c_void_p]
        L.sl_ora_add_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.sl_ora_finalize.argtypes = [C.c_void_p, C.c_void_p]
        L.sl_ora_table.restype = C.c_int64
        L.sl_ora_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.sl_ora_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.POINTER(SealOut),
                                     C.POINTER(SealStats)]
        L.sl_ora_process_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.POINTER(SealOut),
                                        C.POINTER(SealStats), C.c_int32]
        L.sl_ora_scaffold_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _LIB = L
    return _LIB


class SealOracle:
    def __init__(self, cfg: SealCfg):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.sl_ora_create(C.byref(cfg))
        if not self.h:
            raise ValueError("seal oracle: configuration rejected")
        self.n_seqs = 0

    def __del__(self):
        try:
            if self.h:
                self.L.sl_ora_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add_ref(self, bases, offsets):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        self.L.sl_ora_add_ref(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1)
        self.n_seqs += len(offsets) - 1

    def finalize(self):
        v = np.zeros(3, np.int64)
        self.L.sl_ora_finalize(self.h, v.ctypes.data)
        return tuple(int(x) for x in v)

    def table(self):
        n = self.L.sl_ora_table(self.h, None, None, 0)
        keys, ids = np.zeros(n, np.uint64), np.zeros(n, np.int32)
        self.L.sl_ora_table(self.h, keys.ctypes.data, ids.ctypes.data, n)
        return keys, ids

    def process(self, bases, offsets, paired, first_numeric_id=0, threads=1):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        res = SealResult(n_units(self.cfg, n, paired), self.cfg.ids_stride)
        st = SealStats()
        o = res.struct()
        if threads > 1:
            rc = self.L.sl_ora_process_mt(self.h, bases.ctypes.data, offsets.ctypes.data, n, 1 if paired else 0, first_numeric_id,
                                          C.byref(o), C.byref(st), threads)
        else:
            rc = self.L.sl_ora_process(self.h, bases.ctypes.data, offsets.ctypes.data, n, 1 if paired else 0, first_numeric_id,
                                       C.byref(o), C.byref(st))
        if rc:
            raise RuntimeError("seal oracle: process before finalize")
        return res, st

    def scaffold_counts(self):
        n = self.n_seqs + 1
        a = [np.zeros(n, np.int64) for _ in range(4)]
        self.L.sl_ora_scaffold_counts(self.h, *(x.ctypes.data for x in a), n)
        return a


# ---------------------------------------------------------------------------------------------------------
# Independent restatement: plain Python over strings, no rolling registers. Small inputs only.
# ---------------------------------------------------------------------------------------------------------
_CODE = {c: i for i, c in enumerate("ACGT")}
_CODE.update({c.lower(): i for c, i in list(_CODE.items())})
_CODE["U"] = _CODE["u"] = 3


def _f32(x):
    return float(np.float32(x))


class ClosedFormSeal:
    def __init__(self, cfg: SealCfg):
        self.c = cfg
        self.k = k = cfg.k
        self.forbid = bool(cfg.forbid_ns) or cfg.hdist < 1
        self.mm = bool(cfg.mask_middle) or cfg.mid_mask_len > 0
        self.mml = (cfg.mid_mask_len if cfg.mid_mask_len > 0 else 2 - (k & 1)) if self.mm else 0
        if self.mm:
            shift = ((k - self.mml) // 2) * 2
            self.middle = ((1 << 64) - 1) ^ (((1 << (2 * self.mml)) - 1) << shift)
        else:
            self.middle = (1 << 64) - 1
        self.table = {}
        self.seqs = []

    def _rc(self, x):
        k = self.k
        r = 0
        for _ in range(k):
            r = (r << 2) | (3 - (x & 3))
            x >>= 2
        return r

    def _key(self, kmer, rkmer):
        v = max(kmer, rkmer) if self.c.rcomp else kmer
        return (v & self.middle) | (1 << (2 * self.k))

    def _speed_ok(self, key):
        return self.c.speed < 1 or (key & ((1 << 63) - 1)) % 17 >= self.c.speed

    def add_ref(self, seqs):
        self.seqs += list(seqs)

    def finalize(self):
        k, hd = self.k, self.c.hdist
        skip = max(0, self.c.rskip)
        ref_kmers = 0
        for sid, s in enumerate(self.seqs, start=1):
            run = 0
            for i, ch in enumerate(s):
                run = run + 1 if ch in _CODE else 0
                if run < k:
                    continue
                ref_kmers += 1
                if skip > 1 and run % skip != 0:
                    continue
                w = [_CODE[c] for c in s[i - k + 1:i + 1]]
                for nsub in range(hd + 1):
                    for pos in itertools.combinations(range(k), nsub):
                        for alt in itertools.product(range(1, 4), repeat=nsub):
                            v = list(w)
                            for p, a in zip(pos, alt):
                                v[p] = (v[p] + a) & 3
                            x = 0
                            for b in v:
                                x = (x << 2) | b
                            key = self._key(x, self._rc(x))
                            if hd == 0 and not self._speed_ok(key):
                                continue
                            self.table.setdefault(key, set()).add(sid)
        entries = sum(len(v) for v in self.table.values())
        return len(self.table), entries, ref_kmers

    def table_arrays(self):
        items = sorted((k, i) for k, s in self.table.items() for i in s)
        return np.array([k for k, _ in items], np.uint64), np.array([i for _, i in items], np.int32)

    def _hits(self, s, counts):
        """ids of read s appended to the ordered dict `counts` (first-seen order)."""
        k, L = self.k, len(s)
        if L < k or not self.table:
            return
        minlen2 = (k - self.mml) // 2 if self.mm else k
        start = 0 if self.c.restrict_right < 1 else max(0, L - self.c.restrict_right)
        stop = L if self.c.restrict_left < 1 else min(L, self.c.restrict_left)
        for i in range(max(start, k - 1), stop):
            lo = max(start, i - k + 1)
            last_n = -1
            if self.forbid:
                for p in range(i, start - 1, -1):
                    if s[p] == "N":
                        last_n = p
                        break
            ln = i - last_n if last_n >= 0 else i - start + 1
            if ln < minlen2:
                continue
            if self.c.qskip > 1 and i % self.c.qskip != 0:
                continue
            kmer = rkmer = 0
            for j in range(lo, i + 1):
                c = _CODE.get(s[j])
                age = i - j
                if c is not None:
                    kmer |= c << (2 * age)
                    if j > last_n:
                        rkmer |= (3 - c) << (2 * (k - 1 - age))
            key = self._key(kmer, rkmer)
            if not self._speed_ok(key):
                continue
            ids = self.table.get(key)
            if ids:
                for x in sorted(ids):
                    counts[x] = counts.get(x, 0) + 1
                if self.c.match_mode == MATCH_FIRST or (self.c.match_mode == MATCH_UNIQUE and len(ids) == 1):
                    break

    def _valid(self, s):
        run = n = 0
        for ch in s:
            run = run + 1 if ch in _CODE else 0
            n += run >= self.k
        return n

    def _final(self, counts, nvalid):
        mx = max(counts.values()) if counts else 0
        cz = self.c.clearzone
        if self.c.clearzone_fraction > 0:
            cz = max(cz, int(math.ceil(_f32(np.float32(self.c.clearzone_fraction) * np.float32(nvalid)))))
        th = max(1, mx - cz)
        return [i for i, c in counts.items() if c >= th], mx

    def _range(self, fin, numeric_id):
        sites = len(fin)
        if sites < 2 or self.c.ambig_mode == AMBIG_ALL:
            return fin
        if self.c.ambig_mode == AMBIG_TOSS:
            return []
        if self.c.ambig_mode == AMBIG_FIRST:
            return [min(fin)]
        return [fin[numeric_id % sites]]

    def process(self, reads, paired, first_numeric_id=0):
        """reads: list of str (paired: interleaved). Returns per-unit tuples (assigned ids, sites, max), stats, scaffold counters."""
        k = self.k
        n_seq = len(self.seqs) + 1
        sc = np.zeros((4, n_seq), np.int64)
        st = dict(reads_in=0, bases_in=0, reads_matched=0, bases_matched=0, reads_unmatched=0, bases_unmatched=0)
        units = []
        frags = [reads[i:i + 2] for i in range(0, len(reads), 2)] if paired else [[r] for r in reads]
        mkf = np.float32(self.c.min_kmer_fraction)
        for f, mates in enumerate(frags):
            nid = first_numeric_id + f
            st["reads_in"] += len(mates)
            st["bases_in"] += sum(len(m) for m in mates)
            if self.c.keep_pairs_together:
                counts = {}
                for m in mates:
                    self._hits(m, counts)
                fin, mx = self._final(counts, sum(self._valid(m) for m in mates))
                nk = sum(max(len(m) - k + 1, 0) for m in mates)
                minhits = max(self.c.min_kmer_hits, int(_f32(mkf * np.float32(nk))))
                got = []
                rs, ls = len(mates), sum(len(m) for m in mates)
                if mx >= minhits:
                    got = self._range(fin, nid)
                    for i in got:
                        sc[0, i] += rs
                        sc[1, i] += ls
                        sc[2, i] += 1
                        if len(fin) > 1:
                            sc[3, i] += rs
                if got:
                    st["reads_matched"] += rs
                    st["bases_matched"] += ls
                else:
                    st["reads_unmatched"] += rs
                    st["bases_unmatched"] += ls
                units.append((got, len(fin), mx))
            else:
                per = []
                for m in mates:
                    counts = {}
                    self._hits(m, counts)
                    per.append(self._final(counts, self._valid(m)))
                maxes = [p[1] for p in per] + [0]
                for j, (m, (fin, mx)) in enumerate(zip(mates, per)):
                    minhits = max(self.c.min_kmer_hits, int(_f32(mkf * np.float32(max(len(m) - k + 1, 0)))))
                    got = []
                    if mx >= minhits:
                        got = self._range(fin, nid)
                        frag = (maxes[0] >= maxes[1]) if j == 0 else (maxes[1] > maxes[0])
                        for i in got:
                            sc[0, i] += 1
                            sc[1, i] += len(m)
                            sc[2, i] += 1 if frag else 0
                            if len(fin) > 1:
                                sc[3, i] += 1
                        if got:
                            st["reads_matched"] += 1
                            st["bases_matched"] += len(m)
                        else:
                            st["reads_unmatched"] += 1
                            st["bases_unmatched"] += len(m)
                    units.append((got, len(fin), mx))
        return units, st, sc
