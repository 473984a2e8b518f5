#!/usr/bin/env python
"""Full-size parity of config 5 (kmercountexact.sh k=31 on 200 M synthetic 150 bp reads, BASELINE.json): the reads are generated
chunk by chunk on the device and counted there; the same bytes go through the CPU counting oracle (one private table per host
thread over a slice of every chunk -- counting is a sum, the slices' tables are merged at the end); at the end
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
    ap.add_argument("--threads", type=int, default=16, help="host threads = private oracle tables")
    args = ap.parse_args()
    import torch

    from bbtools_b200 import _lib
    from bbtools_b200.kcount import KmerTableSetGPU
    from oracle.kcount import KCountOracle
    lib = _lib.load()
    L = 150
    from concurrent.futures import ThreadPoolExecutor
    tab = KmerTableSetGPU(31, True, initial_keys=1 << 29)
    T = max(1, min(args.threads, os.cpu_count() or 1))
    oras = [KCountOracle(31, True) for _ in range(T)]
    pool = ThreadPoolExecutor(T)
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
        edges = [m * i // T for i in range(T + 1)]
        list(pool.map(lambda i: oras[i].add_reads(hb[edges[i] * L:edges[i + 1] * L], ho[edges[i]:edges[i + 1] + 1] - ho[edges[i]]), range(T)))
        t_cpu += time.perf_counter() - t0
        done += m
    t0 = time.perf_counter()
    ora = oras[0]
    extra_reads = extra_bases = extra_kmers = 0
    for o2 in oras[1:]:  # merge the private tables (counts add, saturating, as the reference's per-thread buffers do)
        n2 = o2.L.kc_ora_dump(o2.h, 1, 0x7FFFFFFF, None, None, 0)
        k2, c2 = np.zeros(max(n2, 1), np.uint64), np.zeros(max(n2, 1), np.int32)
        o2.L.kc_ora_dump(o2.h, 1, 0x7FFFFFFF, k2.ctypes.data, c2.ctypes.data, n2)
        st2 = o2.stats()
        ora.merge_arrays(k2[:n2], c2[:n2])
        extra_reads += st2["reads_in"]
        extra_bases += st2["bases_in"]
        extra_kmers += st2["kmers_in"]
        del k2, c2
        o2.L.kc_ora_destroy(o2.h)
        o2.h = None
    t_merge = time.perf_counter() - t0
    gs, os_ = tab.stats(), ora.stats()
    if T > 1:  # merge_arrays adds keys and counts only; the read / base / k-mer totals of the other slices are added here
        os_["reads_in"] += extra_reads
        os_["bases_in"] += extra_bases
        os_["kmers_in"] += extra_kmers
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
           "gpu_s": round(t_gpu, 3), "oracle_s": round(t_cpu, 3), "oracle_merge_s": round(t_merge, 3), "oracle_threads": T}
    print(json.dumps(res))
    assert res["stats_equal"] and res["khist_equal"] and res["multiset_equal"]


if __name__ == "__main__":
    main()
