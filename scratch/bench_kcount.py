"""quick device-resident timing of the kcount path (cfg 5 shape, scaled)"""
import sys, time, json
sys.path.insert(0, '.')
import torch
from bbtools_b200 import _lib
from bbtools_b200.kcount import KmerTableSetGPU
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 << 20
G = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
L = 150
d_bases = torch.empty(n * L, dtype=torch.uint8, device="cuda")
d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
g = KmerTableSetGPU(31, initial_keys=int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 30)
print(g.table_info())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
for i in range(steps):
    assert lib.kcount_b200_synth_reads(d_bases.data_ptr(), d_off.data_ptr(), n, i * n, L, G, 11, 10, None) == 0
    torch.cuda.synchronize()
    ev[i].record()
    g.add_reads_device(d_bases, d_off, n, n * L)
    ev[i + 1].record() if i == steps - 1 else None
    e2 = torch.cuda.Event(enable_timing=True); e2.record(); torch.cuda.synchronize()
    ms = ev[i].elapsed_time(e2)
    st = g.stats()
    print(json.dumps({"step": i, "ms": ms, "reads_per_s": n / ms * 1e3, "kmers_per_s": n * 120 / ms * 1e3,
                      "GBps_alg(1114B/read)": n * 1114 / ms / 1e6, "unique": st["unique_kmers"], **g.table_info()}))
