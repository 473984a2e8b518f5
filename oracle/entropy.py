"""ctypes binding of the low-entropy filter oracle (oracle/entropy_oracle.c inside libbbduk_oracle.so).

TEST INFRASTRUCTURE ONLY -- never imported by bbtools_b200. PARITY UNPINNED (see entropy_oracle.c)."""
import ctypes as C

import numpy as np

from .oracle import build


class EntropyParams(C.Structure):
    _fields_ = [("cutoff", C.c_float), ("k", C.c_int32), ("window", C.c_int32), ("high_pass", C.c_int32),
                ("remove_pairs_if_either_bad", C.c_int32), ("trim_failures_to_1bp", C.c_int32)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.entropy_ora_values.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.entropy_ora_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(EntropyParams), C.c_void_p]
        L.entropy_ora_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(EntropyParams), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def params(cutoff=0.5, k=5, window=50, high_pass=True, rieb=True, tf1=False) -> EntropyParams:
    p = EntropyParams()
    p.cutoff, p.k, p.window, p.high_pass = cutoff, k, window, int(high_pass)
    p.remove_pairs_if_either_bad, p.trim_failures_to_1bp = int(rieb and not tf1), int(tf1)  # jgi/BBDuk.java:631
    return p


def values(bases, offsets, lo, hi, k=5, window=50):
    bases = np.ascontiguousarray(bases, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    lo = np.ascontiguousarray(lo, np.int32)
    hi = np.ascontiguousarray(hi, np.int32)
    out = np.zeros(len(offsets) - 1, np.float32)
    lib().entropy_ora_values(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, lo.ctypes.data, hi.ctypes.data, k, window,
                             out.ctypes.data)
    return out


def process(bases, offsets, paired, lo, hi, flags, p: EntropyParams):
    """-> (new hi, new flags, [readsEFiltered, basesEFiltered])"""
    bases = np.ascontiguousarray(bases, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    lo = np.ascontiguousarray(lo, np.int32)
    hi2 = np.array(hi, np.int32, copy=True)
    fl2 = np.array(flags, np.uint8, copy=True)
    st = np.zeros(2, np.int64)
    lib().entropy_ora_process(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, int(bool(paired)), lo.ctypes.data,
                              hi2.ctypes.data, fl2.ctypes.data, C.byref(p), st.ctypes.data)
    return hi2, fl2, st


def mask(bases, offsets, paired, lo, hi, flags, p: EntropyParams, mode):
    """entropymask (mode 1: N, 2: lower case) / entropytrim (mode 3) -> (new lo, new hi, mask words, mask_off, [readsEFiltered, basesEFiltered])"""
    bases = np.ascontiguousarray(bases, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    lo2 = np.array(lo, np.int32, copy=True)
    hi2 = np.array(hi, np.int32, copy=True)
    fl = np.ascontiguousarray(flags, np.uint8)
    words = (np.diff(offsets) + 31) // 32
    mask_off = np.zeros(len(offsets), np.int64)
    np.cumsum(words, out=mask_off[1:])
    bits = np.zeros(max(1, int(mask_off[-1])), np.uint32)
    st = np.zeros(2, np.int64)
    lib().entropy_ora_mask(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, int(bool(paired)), lo2.ctypes.data, hi2.ctypes.data,
                           fl.ctypes.data, C.byref(p), int(mode), bits.ctypes.data, mask_off.ctypes.data, st.ctypes.data)
    return lo2, hi2, bits, mask_off, st
