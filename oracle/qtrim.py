"""ctypes binding of the quality-trimming / quality-filter oracle (oracle/qtrim_oracle.c inside libbbduk_oracle.so).

TEST INFRASTRUCTURE ONLY -- never imported by bbtools_b200. PARITY UNPINNED (see qtrim_oracle.c)."""
import ctypes as C

import numpy as np

from .oracle import build


class QtrimParams(C.Structure):
    _fields_ = [("qtrim_left", C.c_int32), ("qtrim_right", C.c_int32), ("trimq", C.c_float), ("min_base_quality", C.c_int32),
                ("max_ns", C.c_int32), ("max_read_length", C.c_int32), ("qual_offset", C.c_int32),
                ("min_read_length", C.c_int32), ("min_len_fraction", C.c_float), ("remove_pairs_if_either_bad", C.c_int32),
                ("trim_failures_to_1bp", C.c_int32), ("trim_poly_a", C.c_int32), ("trim_poly_g_left", C.c_int32),
                ("trim_poly_g_right", C.c_int32), ("filter_poly_g", C.c_int32), ("trim_poly_c_left", C.c_int32),
                ("trim_poly_c_right", C.c_int32), ("filter_poly_c", C.c_int32), ("max_non_poly", C.c_int32),
                ("min_avg_quality", C.c_float), ("min_avg_quality_bases", C.c_int32),
                ("max_n_rate", C.c_float), ("min_consecutive_bases", C.c_int32), ("min_base_frequency", C.c_float),
                ("trim_mode", C.c_int32), ("window_length", C.c_int32), ("min_good_interval", C.c_int32)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.qtrim_ora_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(QtrimParams), C.c_void_p]
        _LIB = L
    return _LIB


def params(qtrim="rl", trimq=6.0, mbq=0, maxns=-1, maxlen=0, qual_offset=33, minlen=10, mlf=0.0, rieb=True, tf1=False,
           polya=0, polyg=(0, 0), fpolyg=0, polyc=(0, 0), fpolyc=0, maxnonpoly=1, maq=0.0, maqb=0, maxnrate=1.0, mcb=0,
           mbf=0.0, mode=0, window=4, goodinterval=2) -> QtrimParams:
    """defaults of jgi/BBDuk.java (:126 trimq, minlen 10, rieb) for the fields the block reads"""
    p = QtrimParams()
    p.qtrim_left, p.qtrim_right = int("l" in qtrim), int("r" in qtrim)
    p.trimq, p.min_base_quality, p.max_ns, p.max_read_length, p.qual_offset = trimq, mbq, maxns, maxlen, qual_offset
    p.min_read_length, p.min_len_fraction = minlen, mlf
    p.remove_pairs_if_either_bad, p.trim_failures_to_1bp = int(rieb and not tf1), int(tf1)  # jgi/BBDuk.java:631
    p.trim_poly_a, p.filter_poly_g, p.filter_poly_c, p.max_non_poly = polya, fpolyg, fpolyc, maxnonpoly
    (p.trim_poly_g_left, p.trim_poly_g_right), (p.trim_poly_c_left, p.trim_poly_c_right) = polyg, polyc
    p.min_avg_quality, p.min_avg_quality_bases = maq, maqb
    p.max_n_rate, p.min_consecutive_bases, p.min_base_frequency = maxnrate, mcb, mbf
    p.trim_mode, p.window_length, p.min_good_interval = mode, window, goodinterval
    if mode == 1:  # qtrim=w trims the right end only (parse/Parser.java:352-357)
        p.qtrim_left, p.qtrim_right = 0, 1
    return p


def process(bases, quals, offsets, paired, lo, hi, flags, p: QtrimParams):
    """-> (new lo, new hi, new flags, stats8)"""
    bases = np.ascontiguousarray(bases, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    lo2 = np.array(lo, np.int32, copy=True)
    hi2 = np.array(hi, np.int32, copy=True)
    fl2 = np.array(flags, np.uint8, copy=True)
    st = np.zeros(8, np.int64)
    q = None if quals is None else np.ascontiguousarray(quals, np.uint8)
    lib().qtrim_ora_process(bases.ctypes.data, None if q is None else q.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                            int(bool(paired)), lo2.ctypes.data, hi2.ctypes.data, fl2.ctypes.data, C.byref(p), st.ctypes.data)
    return lo2, hi2, fl2, st
