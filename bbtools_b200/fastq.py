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


class _BlockReader:
    """a text file (plain or .gz) read in bounded blocks; the bytes of an incomplete last record are carried over"""

    def __init__(self, path, block_bytes):
        self.f = gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")
        self.block = int(block_bytes)
        self.carry = np.zeros(0, np.uint8)
        self.final = False

    def fill(self):
        """append one block to what is held; -> the held text"""
        if not self.final:
            new = np.frombuffer(self.f.read(self.block), np.uint8)
            if new.size < self.block:
                self.final = True  # a short read ends the file (gzip streams included: read() blocks until block or EOF)
            self.carry = np.concatenate([self.carry, new]) if self.carry.size else new
        return self.carry

    def take(self, n_bytes):
        self.carry = self.carry[n_bytes:]

    def close(self):
        self.f.close()


def iter_fastq_blocks(path1, path2=None, block_bytes=256 << 20, unit=1, threads=0):
    """Bounded streaming of one (single / interleaved: unit=2) or two (mates) FASTQ files: yields (text1, text2 or None)
    slices that hold whole records only -- the same number in both files, a multiple of `unit` in a single file -- using
    fastq_b200_index's final=0 / consumed contract. A file that ends inside a record, or mate files of different record
    counts, raise ValueError."""
    lib = _lib.load()
    threads = threads or min(32, os.cpu_count() or 1)
    r1 = _BlockReader(path1, block_bytes)
    r2 = _BlockReader(path2, block_bytes) if path2 else None

    def complete(text, final, max_records=1 << 62):
        """-> (complete records, at most max_records; the bytes they cover)"""
        got, used = C.c_int64(), C.c_int64()
        if text.size == 0:
            return 0, 0
        text = np.ascontiguousarray(text)
        if lib.fastq_b200_index(text.ctypes.data, text.size, int(final), max_records, 1, 0, None, C.byref(got), C.byref(used), threads):
            raise ValueError("malformed FASTQ")
        n = int(got.value)
        if n == 0:
            return 0, 0
        rec = np.zeros(4 * n, np.int64)  # the count-only call does not report the bytes covered
        if lib.fastq_b200_index(text.ctypes.data, text.size, int(final), n, 1, 0, rec.ctypes.data, C.byref(got), C.byref(used), threads) \
                or got.value != n:
            raise ValueError("malformed FASTQ")
        return n, int(used.value)

    try:
        while True:
            t1 = r1.fill()
            t2 = r2.fill() if r2 else None
            n1, u1 = complete(t1, r1.final)
            if r2:
                n2, u2 = complete(t2, r2.final)
                n = min(n1, n2)
                if n < n1:
                    _, u1 = complete(t1, r1.final, n)
                if n < n2:
                    _, u2 = complete(t2, r2.final, n)
            else:
                n = n1 if r1.final else n1 - n1 % unit
                if n < n1:
                    _, u1 = complete(t1, r1.final, n)
            if n > 0:
                yield t1[:u1], (t2[:u2] if r2 else None)
                r1.take(u1)
                if r2:
                    r2.take(u2)
            done = r1.final and (r2 is None or r2.final)
            if done:
                rest = [r.carry for r in (r1, r2) if r is not None]
                if any(np.any((x != 10) & (x != 13)) for x in rest):  # anything but line ends left over
                    raise ValueError("FASTQ input ends inside a record, or the mate files hold different numbers of records")
                return
    finally:
        r1.close()
        if r2:
            r2.close()


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
