"""debug build only: candidate economy of the fast kernel. Build probe_fast.cu with -DBB_FAST_COUNT, link it with the other
objects of bbtools_b200/csrc/build into a library and put that in place of bbtools_b200/libbbduk_b200.so on the GPU box."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bbtools_b200 import _lib  # noqa: E402

lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
wl = bench.WORKLOADS["cfg2"]
eng, stored, _ = bench.build_engine(wl, 0, 0, 1, lib, torch)
L = 150
eng.set_max_read_len(L)
n_pairs = 1 << 20
n_reads = 2 * n_pairs
dev = torch.device("cuda", 0)
outs = {"id0": torch.empty(n_reads, dtype=torch.int32, device=dev), "hi": torch.empty(n_reads, dtype=torch.int32, device=dev),
        "flags": torch.empty(n_reads, dtype=torch.uint8, device=dev)}
d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream(device=dev)
names = ["tiles", "steps", "candidates", "rounds", "forced_before", "forced_after", "enqueue_iters", "tiles_undef",
         "cyc_stageA", "cyc_prepass", "cyc_scanloop", "cyc_rounds_in_loop", "cyc_tails", "cyc_stageD"]
for name, sub, nn in (("standard", 50, 5), ("noN", 50, 0)):
    d_bases = torch.empty(n_reads * L, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
    assert lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), n_pairs, 0, L, C.c_uint64(1), sub, nn, None) == 0
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    raw.bbduk_b200_debug_fast_counters(buf, 1)
    eng.process_device(d_bases, d_off, n_reads, True, outs, d_stats=d_stats, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    raw.bbduk_b200_debug_fast_counters(buf, 1)
    t = buf[0]
    print(name, {n: round(buf[i] / t, 2) for i, n in enumerate(names)}, "per tile of 32 reads; tiles", t, flush=True)
