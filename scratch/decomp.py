"""Where the fast kernel's time goes: the cfg-2 step on variants of the synthetic batch."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bbtools_b200 import _lib  # noqa: E402

lib = _lib.load()
wl = bench.WORKLOADS["cfg2"]
eng, stored, _ = bench.build_engine(wl, 0, 0, 1, lib, torch)
L = 150
eng.set_max_read_len(L)
n_pairs = 4 << 20
n_reads = 2 * n_pairs
dev = torch.device("cuda", 0)
outs = {"id0": torch.empty(n_reads, dtype=torch.int32, device=dev), "hi": torch.empty(n_reads, dtype=torch.int32, device=dev),
        "flags": torch.empty(n_reads, dtype=torch.uint8, device=dev)}
d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream(device=dev)


def timeit(name, d_bases, d_off):
    for _ in range(3):
        eng.process_device(d_bases, d_off, n_reads, True, outs, d_stats=d_stats, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(5):
            eng.process_device(d_bases, d_off, n_reads, True, outs, d_stats=d_stats, stream=stream.cuda_stream)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    trimmed = int((outs["hi"] < L).sum().item())
    print(f"{name:40s} {ms:7.3f} ms  {n_reads / ms / 1e6:7.3f} G reads/s  trimmed {trimmed}", flush=True)


def synth(sub, nn):
    d_bases = torch.empty(n_reads * L, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
    assert lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), n_pairs, 0, L, C.c_uint64(1), sub, nn, None) == 0
    torch.cuda.synchronize()
    return d_bases, d_off


b, o = synth(50, 5)
timeit("cfg2 standard (sub 0.5%, N 0.05%)", b, o)
b1, _ = synth(50, 0)
timeit("no N", b1, o)
b2, _ = synth(0, 0)
timeit("no N, no subs", b2, o)
acgt = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
rnd = acgt[torch.randint(0, 4, (n_reads * L,), device=dev)]
timeit("uniform random ACGT (no adapters, no N)", rnd, o)
# adapters in every read at a fixed position: replace nothing else
allA = torch.full((n_reads * L,), 65, dtype=torch.uint8, device=dev)
timeit("poly-A", allA, o)
