#!/usr/bin/env python
"""Full-size parity run: the whole workload of a BASELINE.json config is generated chunk by chunk on the
device, processed by the CUDA path, and every result record is compared with the multithreaded CPU oracle working on
the same bytes. Nothing is ever written to disk. Prints one JSON summary line.

    python tools/full_parity.py --workload cfg2 --pairs 100000000          # 100 M synthetic 2x150 bp pairs
    python tools/full_parity.py --workload cfg3 --pairs 50000000 --scale 0.02   # cfg-3 shape, 2 Mbp reference
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    ap.add_argument("--pairs", type=int, default=100_000_000, help="pairs (cfg2) / half the reads (cfg3)")
    ap.add_argument("--chunk-pairs", type=int, default=4 << 20)
    ap.add_argument("--scale", type=float, default=1.0, help="cfg3: fraction of the 100 Mbp reference (the oracle's table is host RAM)")
    args = ap.parse_args()
    import torch

    from bbtools_b200 import _lib, make_cfg
    from bbtools_b200.bbduk import BBDukIndexGPU
    from bbtools_b200.fasta import read_fasta
    from oracle.oracle import Oracle
    lib = _lib.load()
    L = 150
    cores = os.cpu_count() or 1
    if args.workload == "cfg2":
        kw = dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
        _, rb, roff = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
        paired, d_ref = True, None
    else:
        kw = dict(k=31)
        n_scaf, scaf_len = 100, int(1_000_000 * args.scale)
        d_ref = torch.empty(n_scaf * scaf_len, dtype=torch.uint8, device="cuda")
        assert lib.bbduk_b200_synth_reference(d_ref.data_ptr(), d_ref.numel(), C.c_uint64(7), None) == 0
        rb = d_ref.cpu().numpy()
        roff = np.arange(0, rb.size + 1, scaf_len, dtype=np.int64)
        paired = False
    gpu, ora = BBDukIndexGPU(make_cfg(**kw)), Oracle(make_cfg(**kw))
    gpu.add_ref(rb, roff)
    ora.add_ref(rb, roff)
    stored = gpu.finalize()
    assert stored == ora.finalize()
    gpu.set_max_read_len(L)
    cp = args.chunk_pairs
    n = 2 * cp
    d_bases = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    outs = {k: torch.empty(n, dtype=torch.int32, device="cuda") for k in ("id0", "lo", "hi", "count")}
    outs["flags"] = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    tot = {}
    crc = 0
    done = 0
    mism = 0
    t_gpu = t_cpu = 0.0
    while done < args.pairs:
        m = min(cp, args.pairs - done)
        nr = 2 * m
        if args.workload == "cfg2":
            rc = lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), m, done, L, C.c_uint64(1), 50, 5, None)
        else:
            rc = lib.bbduk_b200_synth_contam(d_bases.data_ptr(), d_off.data_ptr(), nr, 2 * done, L, d_ref.data_ptr(), d_ref.numel(),
                                             C.c_uint64(1), 10, 100, 5, None)
        assert rc == 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gpu.process_device(d_bases, d_off, nr, paired, outs, d_stats=d_stats)
        torch.cuda.synchronize()
        t_gpu += time.perf_counter() - t0
        hb = d_bases[: nr * L].cpu().numpy()
        ho = np.arange(0, (nr + 1) * L, L, dtype=np.int64)
        t0 = time.perf_counter()
        want, st = ora.process(hb, ho, paired, threads=cores)
        t_cpu += time.perf_counter() - t0
        for name in ("id0", "lo", "hi", "count", "flags"):
            got = outs[name][:nr].cpu().numpy()
            mism += int(np.count_nonzero(got != want.fields()[name]))
            crc = zlib.crc32(got.tobytes(), crc)
        for k_, v in st.as_dict().items():
            tot[k_] = tot.get(k_, 0) + v
        done += m
    dev_tot = dict(zip(tot.keys(), d_stats.cpu().tolist()))
    ro, bo = ora.scaffold_counts()
    rg, bg = gpu.scaffold_counts()
    print(json.dumps({"workload": args.workload, "reads": 2 * args.pairs, "stored_kmers": stored, "mismatching_fields": mism,
                      "counters_equal": dev_tot == tot, "scaffold_counts_equal": bool(np.array_equal(ro, rg) and np.array_equal(bo, bg)),
                      "counters": tot, "crc32_of_results": crc, "gpu_s": round(t_gpu, 3), "oracle_s": round(t_cpu, 3),
                      "oracle_threads": cores}))
    assert mism == 0 and dev_tot == tot


if __name__ == "__main__":
    main()
