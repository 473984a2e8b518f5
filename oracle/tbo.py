"""ctypes binding of the trim-by-overlap oracle (oracle/tbo_oracle.c inside libbbduk_oracle.so).

TEST INFRASTRUCTURE ONLY -- never imported by bbtools_b200. PARITY UNPINNED (see tbo_oracle.c)."""
import ctypes as C

import numpy as np

from .oracle import build


class TboParams(C.Structure):
    _fields_ = [("min_overlap0", C.c_int32), ("min_overlap", C.c_int32), ("min_insert0", C.c_int32), ("min_insert", C.c_int32),
                ("max_ratio", C.c_float), ("min_second_ratio", C.c_float), ("ratio_margin", C.c_float),
                ("ratio_offset", C.c_float), ("g_incr", C.c_float), ("b_incr", C.c_float), ("mee_filter", C.c_float),
                ("qual_offset", C.c_int32)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.tbo_ora_default_params.argtypes = [C.POINTER(TboParams), C.c_int]
        L.tbo_ora_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.tbo_ora_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(TboParams), C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def default_params(strict=True) -> TboParams:
    p = TboParams()
    lib().tbo_ora_default_params(C.byref(p), int(strict))
    return p


def tables():
    comp = np.zeros(128, np.uint8)
    pe = np.zeros(128, np.float32)
    lib().tbo_ora_tables(comp.ctypes.data, pe.ctypes.data)
    return comp, pe


def process(bases, quals, offsets, lo, hi, flags, params=None):
    """-> (new hi, insert per pair, ambig per pair, [reads trimmed, bases trimmed])"""
    p = params or default_params()
    bases = np.ascontiguousarray(bases, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    lo = np.ascontiguousarray(lo, np.int32)
    hi2 = np.array(hi, np.int32, copy=True)
    flags = np.ascontiguousarray(flags, np.uint8)
    n = len(offsets) - 1
    ins = np.full(n // 2, -9, np.int32)
    amb = np.zeros(n // 2, np.uint8)
    st = np.zeros(2, np.int64)
    q = None if quals is None else np.ascontiguousarray(quals, np.uint8)
    lib().tbo_ora_process(bases.ctypes.data, None if q is None else q.ctypes.data, offsets.ctypes.data, n, lo.ctypes.data,
                          hi2.ctypes.data, flags.ctypes.data, C.byref(p), ins.ctypes.data, amb.ctypes.data, st.ctypes.data)
    return hi2, ins, amb, st
