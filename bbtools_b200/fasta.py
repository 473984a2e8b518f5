"""Minimal FASTA/FASTQ readers producing the concatenated (bases, offsets) layout of the C ABI.

Host plumbing only (the reference's stream/ and fileIO/ packages stay out of scope)."""
import gzip

import numpy as np


def _open(path):
    return gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")


def read_fasta(path):
    """-> (names, bases uint8[total], offsets int64[n+1]); sequence lines are joined, case is kept."""
    names, seqs, cur = [], [], None
    with _open(path) as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if not line:
                continue
            if line[:1] == b">":
                if cur is not None:
                    seqs.append(b"".join(cur))
                names.append(line[1:].decode("ascii", "replace"))
                cur = []
            elif cur is not None:
                cur.append(line)
    if cur is not None:
        seqs.append(b"".join(cur))
    return names, *pack(seqs)


def read_fastq(path):
    """-> (names, seqs(list of bytes), quals(list of bytes)); 4-line records."""
    names, seqs, quals = [], [], []
    with _open(path) as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip(b"\r\n")
            f.readline()
            q = f.readline().rstrip(b"\r\n")
            names.append(h.rstrip(b"\r\n")[1:])
            seqs.append(s)
            quals.append(q)
    return names, seqs, quals


def pack(seqs):
    """list of bytes -> (bases uint8[total], offsets int64[n+1])"""
    offsets = np.zeros(len(seqs) + 1, np.int64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
    bases = np.frombuffer(b"".join(seqs), np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return bases, offsets
