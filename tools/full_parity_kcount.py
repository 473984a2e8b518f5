#!/usr/bin/env python
"""Full-size parity of config 5 (kmercountexact.sh k=31 on 200 M synthetic 150 bp reads, BASELINE.json): the reads are generated
chunk by chunk on the device and counted there; the same bytes go through the single-threaded CPU counting oracle; at the end
the two tables are compared as multisets through `Unique Kmers`, kmersIn, the whole count histogram and an order-independent
64-bit checksum over (key, count). Prints one JSON line.

    python tools/full_parity_kcount.py --reads 200000000
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

M1, M2 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F)


def checksum(keys, counts):
    """sum over entries of mix(key) * count, mod 2^64 (order-independent)"""
    acc = np.uint64(0)
    with np.errstate(over="ignore"):
        for a in range(0, len(keys), 1 << 26):
            k = keys[a:a + (1 << 26)].astype(np.uint64)
            x = (k ^ (k >> np.uint64(31))) * M1
            x = (x ^ (x >> np.uint64(29))) * M2
            acc += (x * counts[a:a + (1 << 26)].astype(np.uint64)).sum(dtype=np.uint64)
    return int(acc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=200_000_000)
    ap.add_argument("--chunk-reads", type=int, default=8 << 20)
    ap.add_argument("--genome", type=int, default=100_000_000)
    args = ap.parse_args()
    import torch

    from bbtools_b200 import _lib
    from bbtools_b200.kcount import KmerTableSetGPU
    from oracle.kcount import KCountOracle
    lib = _lib.load()
    L = 150
    tab = KmerTableSetGPU(31, True, initial_keys=1 << 29)
    ora = KCountOracle(31, True)
    n = args.chunk_reads
    d_bases = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    done, t_gpu, t_cpu = 0, 0.0, 0.0
    while done < args.reads:
        m = min(n, args.reads - done)
        assert lib.kcount_b200_synth_reads(d_bases.data_ptr(), d_off.data_ptr(), m, done, L, args.genome, C.c_uint64(11), 10, None) == 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tab.add_reads_device(d_bases, d_off, m, m * L)
        torch.cuda.synchronize()
        t_gpu += time.perf_counter() - t0
        hb = d_bases[: m * L].cpu().numpy()
        ho = np.arange(0, (m + 1) * L, L, dtype=np.int64)
        t0 = time.perf_counter()
        ora.add_reads(hb, ho)
        t_cpu += time.perf_counter() - t0
        done += m
    gs, os_ = tab.stats(), ora.stats()
    gh, oh = tab.khist(100000), ora.khist(100000)
    gk, gc = tab.dump()
    g_sum = checksum(gk, gc)
    n_g = len(gk)
    del gk, gc
    n_o = ora.L.kc_ora_dump(ora.h, 1, 0x7FFFFFFF, None, None, 0)  # unsorted: the checksum does not care
    ok, oc = np.zeros(max(n_o, 1), np.uint64), np.zeros(max(n_o, 1), np.int32)
    ora.L.kc_ora_dump(ora.h, 1, 0x7FFFFFFF, ok.ctypes.data, oc.ctypes.data, n_o)
    ok, oc = ok[:n_o], oc[:n_o]
    o_sum = checksum(ok, oc)
    res = {"workload": "cfg5", "reads": args.reads, "gpu_stats": gs, "oracle_stats": os_, "stats_equal": gs == os_,
           "khist_equal": bool(np.array_equal(np.asarray(gh), np.asarray(oh))), "entries_gpu": n_g, "entries_oracle": len(ok),
           "checksum_gpu": g_sum, "checksum_oracle": o_sum, "multiset_equal": g_sum == o_sum and n_g == len(ok),
           "gpu_s": round(t_gpu, 3), "oracle_s": round(t_cpu, 3), "oracle_threads": 1}
    print(json.dumps(res))
    assert res["stats_equal"] and res["khist_equal"] and res["multiset_equal"]


if __name__ == "__main__":
    main()
