"""Seal's k-mer matching path on B200 (include/seal_b200.h): ctypes mirror of the POD structs and the host-side class.

Mirrors jgi.Seal's loader + matching block (jgi/Seal.java:1760-1946, :2186-2276). No CPU fallback: the CUDA library
must be built (bbtools_b200/_lib.py raises ImportError otherwise)."""
import ctypes as C

import numpy as np

AMBIG_ALL, AMBIG_FIRST, AMBIG_TOSS, AMBIG_RANDOM = 1, 2, 3, 4
MATCH_ALL, MATCH_FIRST, MATCH_UNIQUE = 1, 2, 3


class SealCfg(C.Structure):
    """include/seal_b200.h seal_cfg (keep the field order in sync with the header)."""
    _fields_ = [
        ("struct_size", C.c_int32),
        ("k", C.c_int32),
        ("rcomp", C.c_int32),
        ("mask_middle", C.c_int32),
        ("mid_mask_len", C.c_int32),
        ("forbid_ns", C.c_int32),
        ("hdist", C.c_int32),
        ("speed", C.c_int32),
        ("qskip", C.c_int32),
        ("rskip", C.c_int32),
        ("restrict_left", C.c_int32),
        ("restrict_right", C.c_int32),
        ("ambig_mode", C.c_int32),
        ("match_mode", C.c_int32),
        ("keep_pairs_together", C.c_int32),
        ("clearzone", C.c_int32),
        ("clearzone_fraction", C.c_float),
        ("min_kmer_hits", C.c_int32),
        ("min_kmer_fraction", C.c_float),
        ("device", C.c_int32),
        ("table_load_pct", C.c_int32),
        ("ids_stride", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


class SealOut(C.Structure):
    _fields_ = [("n_assigned", C.c_void_p), ("first_id", C.c_void_p), ("n_sites", C.c_void_p), ("max_hits", C.c_void_p),
                ("ids", C.c_void_p)]


class SealStats(C.Structure):
    _fields_ = [("reads_in", C.c_int64), ("bases_in", C.c_int64), ("reads_matched", C.c_int64), ("bases_matched", C.c_int64),
                ("reads_unmatched", C.c_int64), ("bases_unmatched", C.c_int64), ("reserved", C.c_int64 * 2)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


def make_cfg(**kw) -> SealCfg:
    """Seal's defaults (jgi/Seal.java:104-128, :3088-3098) with overrides."""
    c = SealCfg()
    c.struct_size = C.sizeof(SealCfg)
    c.k, c.rcomp, c.mask_middle, c.mid_mask_len, c.forbid_ns, c.hdist = 31, 1, 1, 0, 0, 0
    c.ambig_mode, c.match_mode, c.keep_pairs_together = AMBIG_RANDOM, MATCH_ALL, 1
    c.min_kmer_hits, c.table_load_pct, c.ids_stride = 1, 50, 4
    for k, v in kw.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


class SealResult:
    def __init__(self, n_units, stride):
        self.n_assigned = np.zeros(n_units, np.int32)
        self.first_id = np.zeros(n_units, np.int32)
        self.n_sites = np.zeros(n_units, np.int32)
        self.max_hits = np.zeros(n_units, np.int32)
        self.ids = np.zeros(max(1, n_units * max(stride, 0)), np.int32)
        self.stride = stride

    def struct(self):
        o = SealOut()
        o.n_assigned, o.first_id = self.n_assigned.ctypes.data, self.first_id.ctypes.data
        o.n_sites, o.max_hits = self.n_sites.ctypes.data, self.max_hits.ctypes.data
        o.ids = self.ids.ctypes.data if self.stride > 0 else None
        return o

    def fields(self):
        return {"n_assigned": self.n_assigned, "first_id": self.first_id, "n_sites": self.n_sites, "max_hits": self.max_hits,
                "ids": self.ids}


def n_units(cfg, n_reads, paired):
    return n_reads // 2 if (paired and cfg.keep_pairs_together) else n_reads



class SealIndexGPU:
    """Reference table + matcher on one GPU: add_ref -> finalize -> process (host numpy buffers)."""

    def __init__(self, cfg: SealCfg):
        from . import _lib
        self.lib = _lib.load()
        self.cfg = cfg
        h = C.c_void_p()
        if self.lib.seal_b200_create(C.byref(cfg), C.byref(h)):
            raise ValueError(self.lib.seal_b200_last_error(None).decode())
        self.h = h
        self.n_seqs = 0

    def _check(self, rc):
        if rc:
            raise RuntimeError(self.lib.seal_b200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.seal_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_ref(self, bases, offsets):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        self._check(self.lib.seal_b200_add_ref(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1))
        self.n_seqs += len(offsets) - 1

    def finalize(self):
        """(storedKmers, (k-mer, id) entries, refKmers)"""
        v = np.zeros(3, np.int64)
        self._check(self.lib.seal_b200_finalize(self.h, v.ctypes.data))
        return tuple(int(x) for x in v)

    def table(self):
        n = C.c_int64()
        self._check(self.lib.seal_b200_table_export(self.h, None, None, 0, C.byref(n)))
        keys, ids = np.zeros(n.value, np.uint64), np.zeros(n.value, np.int32)
        self._check(self.lib.seal_b200_table_export(self.h, keys.ctypes.data, ids.ctypes.data, n.value, C.byref(n)))
        return keys, ids

    def process(self, bases, offsets, paired, first_numeric_id=0):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        nu = self.lib.seal_b200_n_units(self.h, n, 1 if paired else 0)
        res = SealResult(nu, self.cfg.ids_stride)
        st = SealStats()
        o = res.struct()
        self._check(self.lib.seal_b200_process(self.h, bases.ctypes.data, offsets.ctypes.data, n, 1 if paired else 0,
                                               first_numeric_id, C.byref(o), C.byref(st)))
        return res, st

    def scaffold_counts(self):
        n = self.n_seqs + 1
        a = [np.zeros(n, np.int64) for _ in range(4)]
        self._check(self.lib.seal_b200_scaffold_counts(self.h, *(x.ctypes.data for x in a), n))
        return a

    @property
    def launches(self):
        return int(self.lib.seal_b200_launch_count(self.h))
