"""ctypes binding of the KmerCountExact oracle (oracle/kcount_oracle.c inside libbbduk_oracle.so).

TEST INFRASTRUCTURE ONLY -- never imported by bbtools_b200. PARITY UNPINNED (see kcount_oracle.c)."""
import ctypes as C

import numpy as np

from .oracle import build

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.kc_ora_create.restype = C.c_void_p
        L.kc_ora_create.argtypes = [C.c_int, C.c_int]
        L.kc_ora_destroy.argtypes = [C.c_void_p]
        L.kc_ora_add_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.kc_ora_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.kc_ora_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.kc_ora_khist.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.kc_ora_dump.restype = C.c_int64
        L.kc_ora_dump.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
        _LIB = L
    return _LIB


class KCountOracle:
    def __init__(self, k=31, rcomp=True):
        self.L = lib()
        self.k, self.rcomp = k, bool(rcomp)
        self.h = self.L.kc_ora_create(k, 1 if rcomp else 0)
        if not self.h:
            raise ValueError("k must be in [1,31]")

    def __del__(self):
        try:
            if self.h:
                self.L.kc_ora_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add_reads(self, bases, offsets):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        self.L.kc_ora_add_reads(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1)

    def merge_arrays(self, keys, counts):
        keys = np.ascontiguousarray(keys, np.uint64)
        counts = np.ascontiguousarray(counts, np.int32)
        self.L.kc_ora_merge(self.h, keys.ctypes.data, counts.ctypes.data, len(keys))

    def stats(self):
        v = np.zeros(4, np.int64)
        self.L.kc_ora_stats(self.h, v.ctypes.data)
        return {"reads_in": int(v[0]), "bases_in": int(v[1]), "kmers_in": int(v[2]), "unique_kmers": int(v[3])}

    def khist(self, histmax=100000):
        hist = np.zeros(histmax + 1, np.int64)
        self.L.kc_ora_khist(self.h, histmax, hist.ctypes.data)
        return hist

    def dump(self, mincount=1, maxcount=0x7FFFFFFF):
        n = self.L.kc_ora_dump(self.h, mincount, maxcount, None, None, 0)
        keys = np.zeros(max(n, 1), np.uint64)
        counts = np.zeros(max(n, 1), np.int32)
        self.L.kc_ora_dump(self.h, mincount, maxcount, keys.ctypes.data, counts.ctypes.data, n)
        order = np.argsort(keys[:n])
        return keys[:n][order], counts[:n][order]
