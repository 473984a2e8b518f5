"""Native FASTQ feed (include/fastq_b200.h): text buffers <-> the batch layout of bbduk_b200_process.

Host plumbing of SURVEY.md 8f row 1: replaces the quad-line record splitting of stream/FastqStreamer.java /
stream/FASTQ.java and the record formatting of stream/ReadStreamByteWriter.java for the reads the k-mer block keeps."""
import ctypes as C
import gzip
import os

import numpy as np

from . import _lib


def load_text(path) -> np.ndarray:
    """whole file as a uint8 array (gzip is inflated by Python: decompression is not part of this path)"""
    if str(path).endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return np.frombuffer(f.read(), np.uint8)
    return np.fromfile(path, np.uint8)


class FastqBatch:
    """records of one (single/interleaved) or two (mates) FASTQ texts, indexed natively"""

    def __init__(self, text1: np.ndarray, text2: np.ndarray = None, threads: int = 0):
        self.lib = _lib.load()
        self.threads = threads or min(32, os.cpu_count() or 1)
        self.text1 = np.ascontiguousarray(text1, np.uint8)
        self.text2 = None if text2 is None else np.ascontiguousarray(text2, np.uint8)
        stride = 2 if self.text2 is not None else 1
        n1 = self._count(self.text1)
        n2 = self._count(self.text2) if self.text2 is not None else n1
        if self.text2 is not None and n1 != n2:
            raise ValueError(f"the two files hold different numbers of records ({n1} vs {n2})")
        self.n_reads = n1 * stride
        self.rec = np.zeros(4 * max(self.n_reads, 1), np.int64)
        self._index(self.text1, n1, stride, 0)
        if self.text2 is not None:
            self._index(self.text2, n2, stride, 1)

    def _count(self, text):
        # 4-line records: a final line without '\n' still counts
        if text.size == 0:
            return 0
        got, used = C.c_int64(), C.c_int64()
        if self.lib.fastq_b200_index(text.ctypes.data, text.size, 1, 1 << 62, 1, 0, None, C.byref(got), C.byref(used), self.threads):
            raise ValueError("malformed FASTQ")
        return int(got.value)

    def _index(self, text, n, stride, first):
        if n == 0:
            return
        got, used = C.c_int64(), C.c_int64()
        rc = self.lib.fastq_b200_index(text.ctypes.data, text.size, 1, n, stride, first, self.rec.ctypes.data,
                                       C.byref(got), C.byref(used), self.threads)
        if rc or got.value != n:
            raise ValueError(f"malformed FASTQ (rc={rc}, {got.value} of {n} records indexed)")

    def _t2(self):
        return None if self.text2 is None else self.text2.ctypes.data

    def arrays(self):
        """-> (bases uint8[total], offsets int64[n+1]) in the layout of bbduk_b200_process"""
        offsets = np.zeros(self.n_reads + 1, np.int64)
        self.lib.fastq_b200_gather(self.text1.ctypes.data, self._t2(), self.rec.ctypes.data, self.n_reads, None,
                                   offsets.ctypes.data, self.threads)
        bases = np.empty(int(offsets[-1]), np.uint8)
        rc = self.lib.fastq_b200_gather(self.text1.ctypes.data, self._t2(), self.rec.ctypes.data, self.n_reads,
                                        bases.ctypes.data, offsets.ctypes.data, self.threads)
        if rc:
            raise RuntimeError("fastq_b200_gather failed")
        return bases, offsets

    def quals(self):
        """quality bytes in the layout of arrays()[0] (the same gather with the records pointed at their quality lines)"""
        rq = self.rec.copy()
        rq[1::4] = self.rec[3::4]
        offsets = np.zeros(self.n_reads + 1, np.int64)
        self.lib.fastq_b200_gather(self.text1.ctypes.data, self._t2(), rq.ctypes.data, self.n_reads, None,
                                   offsets.ctypes.data, self.threads)
        q = np.empty(int(offsets[-1]), np.uint8)
        if self.lib.fastq_b200_gather(self.text1.ctypes.data, self._t2(), rq.ctypes.data, self.n_reads, q.ctypes.data,
                                      offsets.ctypes.data, self.threads):
            raise RuntimeError("fastq_b200_gather failed")
        return q

    def format(self, per, lo, hi, flags, removed=False, mate_sel=0, trim_removed=False) -> np.ndarray:
        """FASTQ text of the kept (or removed) units, trimmed to [lo,hi)"""
        lo = np.ascontiguousarray(lo, np.int32)
        hi = np.ascontiguousarray(hi, np.int32)
        flags = np.ascontiguousarray(flags, np.uint8)
        n = C.c_int64()
        args = (self.text1.ctypes.data, self._t2(), self.rec.ctypes.data, self.n_reads, per, lo.ctypes.data, hi.ctypes.data,
                flags.ctypes.data, int(removed), mate_sel, int(trim_removed))
        if self.lib.fastq_b200_format(*args, None, 0, C.byref(n), self.threads):
            raise RuntimeError("fastq_b200_format failed")
        out = np.empty(n.value, np.uint8)
        if self.lib.fastq_b200_format(*args, out.ctypes.data, out.size, C.byref(n), self.threads):
            raise RuntimeError("fastq_b200_format failed")
        return out
