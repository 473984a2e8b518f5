#!/usr/bin/env python
"""Throughput of the native FASTQ feed (host only): index + gather of N synthetic 2x150 bp pairs and formatting of the
trimmed output, against the plain-Python reader. Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bbtools_b200 import synth  # noqa: E402
from bbtools_b200.fastq import FastqBatch  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
bases, offsets = synth.paired_adapter_reads(n_pairs, seed=1)
L = 150
reads = bases.reshape(-1, L)
rec = np.empty((2 * n_pairs, 22 + L + 1 + 2 + L + 1), np.uint8)
hdr = np.frombuffer(b"@r%019d\n" % 0, np.uint8)
rec[:, :22] = hdr
idx = np.arange(2 * n_pairs)
for d in range(19):
    rec[:, 21 - 1 - d] = 48 + (idx // 10 ** d) % 10
rec[:, 22:22 + L] = reads
rec[:, 22 + L] = 10
rec[:, 22 + L + 1:22 + L + 3] = np.frombuffer(b"+\n", np.uint8)
rec[:, 22 + L + 3:22 + 2 * L + 3] = 73
rec[:, -1] = 10
rec[:, 21] = 10
text = rec.reshape(-1)
t0 = time.perf_counter()
fb = FastqBatch(text, threads=threads)
t_index = time.perf_counter() - t0
t0 = time.perf_counter()
b2, o2 = fb.arrays()
t_gather = time.perf_counter() - t0
assert np.array_equal(b2, bases) and np.array_equal(o2, offsets)
lo = np.zeros(2 * n_pairs, np.int32)
hi = np.full(2 * n_pairs, 120, np.int32)
flags = np.zeros(2 * n_pairs, np.uint8)
t0 = time.perf_counter()
out = fb.format(2, lo, hi, flags)
t_fmt = time.perf_counter() - t0
print(json.dumps({"reads": 2 * n_pairs, "text_bytes": int(text.size), "threads": threads,
                  "index_GBps": text.size / t_index / 1e9, "gather_GBps": text.size / t_gather / 1e9,
                  "parse_reads_per_s": 2 * n_pairs / (t_index + t_gather), "format_GBps": out.size / t_fmt / 1e9,
                  "format_reads_per_s": 2 * n_pairs / t_fmt}))
