"""ctypes mirror of include/bbduk_b200.h (POD structs only; keep field order in sync with the header)."""
import ctypes as C

import numpy as np

ABI_VERSION = 2
GEN_JGI = 0
GEN_S = 1

F_DISCARDED = 0x01
F_REMOVED = 0x02
F_KTRIMMED = 0x04
F_TPE = 0x08
F_SPLIT = 0x10
F_TBO = 0x20


class BBDukCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("generation", C.c_int32),
        ("k", C.c_int32),
        ("mink", C.c_int32),
        ("use_short_kmers", C.c_int32),
        ("hdist", C.c_int32),
        ("hdist2", C.c_int32),
        ("edist", C.c_int32),
        ("edist2", C.c_int32),
        ("qhdist", C.c_int32),
        ("qhdist2", C.c_int32),
        ("rcomp", C.c_int32),
        ("mask_middle", C.c_int32),
        ("mid_mask_len", C.c_int32),
        ("forbid_ns", C.c_int32),
        ("ktrim_left", C.c_int32),
        ("ktrim_right", C.c_int32),
        ("ktrim_n", C.c_int32),
        ("ksplit", C.c_int32),
        ("ktrim_exclusive", C.c_int32),
        ("trim_pad", C.c_int32),
        ("restrict_left", C.c_int32),
        ("restrict_right", C.c_int32),
        ("skip_r1", C.c_int32),
        ("skip_r2", C.c_int32),
        ("qskip", C.c_int32),
        ("speed", C.c_int32),
        ("min_skip", C.c_int32),
        ("max_skip", C.c_int32),
        ("max_bad_kmers", C.c_int32),
        ("min_kmer_fraction", C.c_float),
        ("min_covered_fraction", C.c_float),
        ("find_best_match", C.c_int32),
        ("kmask_fully_covered", C.c_int32),
        ("kmask_lowercase", C.c_int32),
        ("trim_symbol", C.c_int32),
        ("min_read_length", C.c_int32),
        ("min_len_fraction", C.c_float),
        ("require_both_bad", C.c_int32),
        ("trim_pairs_evenly", C.c_int32),
        ("trim_failures_to_1bp", C.c_int32),
        ("device", C.c_int32),
        ("table_load_pct", C.c_int32),
        ("minlen2", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


class BBDukOut(C.Structure):
    _fields_ = [
        ("id0", C.c_void_p),
        ("id0b", C.c_void_p),
        ("lo", C.c_void_p),
        ("hi", C.c_void_p),
        ("flags", C.c_void_p),
        ("count", C.c_void_p),
        ("maskbits", C.c_void_p),
        ("mask_off", C.c_void_p),
    ]


class BBDukTboCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("strict_overlap", C.c_int32),
        ("min_overlap0", C.c_int32),
        ("min_overlap", C.c_int32),
        ("min_insert0", C.c_int32),
        ("min_insert", C.c_int32),
        ("qual_offset", C.c_int32),
        ("mee_filter", C.c_float),
        ("reserved", C.c_int32 * 4),
    ]


class BBDukQtrimCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("qtrim_left", C.c_int32),
        ("qtrim_right", C.c_int32),
        ("trimq", C.c_float),
        ("min_base_quality", C.c_int32),
        ("max_ns", C.c_int32),
        ("max_read_length", C.c_int32),
        ("qual_offset", C.c_int32),
        ("trim_poly_a", C.c_int32),
        ("trim_poly_g_left", C.c_int32),
        ("trim_poly_g_right", C.c_int32),
        ("filter_poly_g", C.c_int32),
        ("trim_poly_c_left", C.c_int32),
        ("trim_poly_c_right", C.c_int32),
        ("filter_poly_c", C.c_int32),
        ("max_non_poly", C.c_int32),
        ("min_avg_quality", C.c_float),
        ("min_avg_quality_bases", C.c_int32),
        ("max_n_rate", C.c_float),
        ("min_consecutive_bases", C.c_int32),
        ("min_base_frequency", C.c_float),
        ("trim_mode", C.c_int32),
        ("window_length", C.c_int32),
        ("min_good_interval", C.c_int32),
    ]


class BBDukEntropyCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("cutoff", C.c_float),
        ("k", C.c_int32),
        ("window", C.c_int32),
        ("high_pass", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class BBDukChainCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("do_tbo", C.c_int32),
        ("do_qtrim", C.c_int32),
        ("do_entropy", C.c_int32),
        ("tbo", BBDukTboCfg),
        ("qtrim", BBDukQtrimCfg),
        ("entropy", BBDukEntropyCfg),
    ]


class BBDukStats(C.Structure):
    _fields_ = [
        ("reads_in", C.c_int64),
        ("bases_in", C.c_int64),
        ("reads_ktrimmed", C.c_int64),
        ("bases_ktrimmed", C.c_int64),
        ("reads_kfiltered", C.c_int64),
        ("bases_kfiltered", C.c_int64),
        ("reads_out", C.c_int64),
        ("bases_out", C.c_int64),
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class BBDukTableDesc(C.Structure):
    _fields_ = [
        ("n_slots", C.c_int64),
        ("n_filter_words", C.c_int64),
        ("stored_kmers", C.c_int64),
        ("n_scaffolds", C.c_int32),
        ("reserved", C.c_int32),
        ("d_keys", C.c_void_p),
        ("d_vals", C.c_void_p),
        ("d_filter", C.c_void_p),
        ("scalars", C.c_int64 * 8),
    ]


def default_cfg() -> BBDukCfg:
    """Defaults of jgi.BBDuk's constructor (reference jgi/BBDuk.java:107-147, :4953-4977)."""
    c = BBDukCfg()
    c.struct_size = C.sizeof(BBDukCfg)
    c.generation = GEN_JGI
    c.k = 0
    c.mink = -1
    c.hdist2 = c.edist2 = c.qhdist2 = -1
    c.rcomp = 1
    c.mask_middle = 1
    c.qskip = 1
    c.min_skip = c.max_skip = 1
    c.trim_symbol = ord("N")
    c.min_read_length = 10
    c.device = -1
    return c


def make_cfg(**kw) -> BBDukCfg:
    c = default_cfg()
    names = {n for n, _ in BBDukCfg._fields_}
    for key, val in kw.items():
        if key not in names:
            raise KeyError(f"unknown bbduk_cfg field {key!r}")
        setattr(c, key, val)
    return c


class Outputs:
    """Host-side struct-of-arrays result of one process call."""

    def __init__(self, n_reads: int, lengths=None, want_mask: bool = False):
        self.n = n_reads
        self.id0 = np.full(n_reads, -7, np.int32)
        self.id0b = np.full(n_reads, -7, np.int32)
        self.lo = np.full(n_reads, -7, np.int32)
        self.hi = np.full(n_reads, -7, np.int32)
        self.flags = np.full(n_reads, 0xEE, np.uint8)
        self.count = np.full(n_reads, -7, np.int32)
        self.mask_off = None
        self.maskbits = None
        if want_mask:
            words = (np.asarray(lengths, np.int64) + 31) // 32
            self.mask_off = np.zeros(n_reads + 1, np.int64)
            np.cumsum(words, out=self.mask_off[1:])
            self.maskbits = np.full(int(self.mask_off[-1]), 0xDEADBEEF, np.uint32)

    def struct(self) -> BBDukOut:
        o = BBDukOut()
        for name in ("id0", "id0b", "lo", "hi", "flags", "count", "maskbits", "mask_off"):
            a = getattr(self, name)
            setattr(o, name, a.ctypes.data if a is not None and a.size else None)
        return o

    def fields(self):
        d = {n: getattr(self, n) for n in ("id0", "id0b", "lo", "hi", "flags", "count")}
        if self.maskbits is not None:
            d["maskbits"] = self.maskbits
        return d
