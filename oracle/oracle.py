"""ctypes binding of oracle/libbbduk_oracle.so (the C restatement of the reference's BBDuk k-mer path).

TEST INFRASTRUCTURE ONLY -- never imported by bbtools_b200. PARITY UNPINNED (no reference golden
vectors exist and no JVM is available; see oracle/bbduk_oracle.c header)."""
import ctypes as C
import os
import subprocess

import numpy as np

from bbtools_b200._abi import BBDukCfg, BBDukOut, BBDukStats, Outputs

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libbbduk_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("bbduk_oracle.c", "kcount_oracle.c", "tbo_oracle.c", "qtrim_oracle.c", "entropy_oracle.c", "seal_oracle.c", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "bbduk_b200.h"))
    srcs.append(os.path.join(_HERE, "..", "include", "seal_b200.h"))
    stale = (not os.path.exists(so)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ora_create.restype = C.c_void_p
        L.ora_create.argtypes = [C.POINTER(BBDukCfg)]
        L.ora_error.restype = C.c_char_p
        L.ora_destroy.argtypes = [C.c_void_p]
        L.ora_add_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.ora_finalize.restype = C.c_int64
        L.ora_finalize.argtypes = [C.c_void_p]
        L.ora_n_scaffolds.restype = C.c_int32
        L.ora_n_scaffolds.argtypes = [C.c_void_p]
        L.ora_dump_table.restype = C.c_int64
        L.ora_dump_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.ora_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                  C.POINTER(BBDukOut), C.POINTER(BBDukStats), C.c_int]
        L.ora_scaffold_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.ora_derived.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_lookup_key.restype = C.c_int32
        L.ora_lookup_key.argtypes = [C.c_void_p, C.c_uint64]
        _LIB = L
    return _LIB


class Oracle:
    def __init__(self, cfg: BBDukCfg):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.ora_create(C.byref(cfg))
        if not self.h:
            raise ValueError(self.L.ora_error().decode())

    def close(self):
        if self.h:
            self.L.ora_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_ref(self, bases: np.ndarray, offsets: np.ndarray):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        rc = self.L.ora_add_ref(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1)
        if rc:
            raise RuntimeError("ora_add_ref failed")

    def finalize(self) -> int:
        return int(self.L.ora_finalize(self.h))

    @property
    def n_scaffolds(self) -> int:
        return int(self.L.ora_n_scaffolds(self.h))

    def dump_table(self):
        n = self.L.ora_dump_table(self.h, None, None, 0)
        keys = np.zeros(n, np.uint64)
        vals = np.zeros(n, np.int32)
        self.L.ora_dump_table(self.h, keys.ctypes.data, vals.ctypes.data, n)
        order = np.argsort(keys)
        return keys[order], vals[order]

    def derived(self):
        v = np.zeros(16, np.int64)
        self.L.ora_derived(self.h, v.ctypes.data)
        names = ["k", "kbig", "mink", "useShortKmers", "maskMiddle", "midMaskLen", "minlen", "minlen2", "minminlen",
                 "forbidNs", "hdist", "hdist2", "middleMask", "mask", "kfilter", "rieb"]
        return dict(zip(names, (int(x) for x in v)))

    def lookup_key(self, key: int) -> int:
        return int(self.L.ora_lookup_key(self.h, C.c_uint64(key)))

    def process(self, bases: np.ndarray, offsets: np.ndarray, paired: bool, threads: int = 1, want_mask=False):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        out = Outputs(n, np.diff(offsets), want_mask=want_mask)
        st = BBDukStats()
        o = out.struct()
        rc = self.L.ora_process(self.h, bases.ctypes.data, offsets.ctypes.data, n, int(bool(paired)),
                                C.byref(o), C.byref(st), threads)
        if rc:
            raise RuntimeError("ora_process failed")
        return out, st

    def scaffold_counts(self):
        n = self.n_scaffolds + 1
        rc_ = np.zeros(n, np.int64)
        bc = np.zeros(n, np.int64)
        self.L.ora_scaffold_counts(self.h, rc_.ctypes.data, bc.ctypes.data, n)
        return rc_, bc
