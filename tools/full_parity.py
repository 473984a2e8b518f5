#!/usr/bin/env python
"""Full-size parity run: the whole workload of a BASELINE.json config is generated chunk by chunk on the
device, processed by the CUDA path, and every result record is compared with the multithreaded CPU oracle working on
the same bytes. Nothing is ever written to disk. Prints one JSON summary line.

    python tools/full_parity.py --workload cfg2 --pairs 100000000          # 100 M synthetic 2x150 bp pairs
    python tools/full_parity.py --workload cfg3 --pairs 250000000                # 500 M reads vs the 100 Mbp reference
    python tools/full_parity.py --workload cfg3 --pairs 50000000 --scale 0.02   # cfg-3 shape, 2 Mbp reference
    python tools/full_parity.py --workload cfg4 --pairs 100000000               # k=27 hdist=2 vs 100 x 1 kbp (2.9e8 keys)
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
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4"])
    ap.add_argument("--pairs", type=int, default=100_000_000, help="pairs (cfg2) / half the reads (cfg3)")
    ap.add_argument("--chunk-pairs", type=int, default=4 << 20)
    ap.add_argument("--scale", type=float, default=1.0, help="cfg3: fraction of the 100 Mbp reference (the oracle's table is host RAM)")
    ap.add_argument("--mode", default="ktrimr", choices=["ktrimr", "ktriml", "kfilter", "kmask"],
                    help="cfg2 reads and adapters through another mode of the tuned kernel (ktrim=l, kfilter, ktrim=N)")
    ap.add_argument("--tbo", action="store_true", help="cfg2: follow the k-mer block with trim-by-overlap (the `tpe tbo` command)")
    ap.add_argument("--qtrim", action="store_true", help="cfg2: then quality trimming qtrim=rl trimq=10 on synthetic decaying qualities")
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
        kw = {"ktrimr": dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1),
              "ktriml": dict(k=23, mink=11, hdist=1, ktrim_left=1),
              "kfilter": dict(k=23, hdist=1),
              "kmask": dict(k=23, mink=11, hdist=1, ktrim_n=1)}[args.mode]
        _, rb, roff = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
        paired, d_ref = True, None
    elif args.workload == "cfg4":
        kw = dict(k=27, hdist=2)
        n_scaf, scaf_len = 100, 1000
        d_ref = torch.empty(n_scaf * scaf_len, dtype=torch.uint8, device="cuda")
        assert lib.bbduk_b200_synth_reference(d_ref.data_ptr(), d_ref.numel(), C.c_uint64(9), None) == 0
        rb = d_ref.cpu().numpy()
        roff = np.arange(0, rb.size + 1, scaf_len, dtype=np.int64)
        paired = True
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
    want_mask = args.workload == "cfg2" and args.mode == "kmask"
    if want_mask:
        words = (L + 31) // 32
        outs["mask_off"] = torch.arange(0, (n + 1) * words, words, dtype=torch.int64, device="cuda")
        outs["maskbits"] = torch.zeros(n * words, dtype=torch.int32, device="cuda")
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(cores)
    tbo_tot, q_tot = np.zeros(2, np.int64), np.zeros(8, np.int64)
    d_tst = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_qst = torch.zeros(8, dtype=torch.int64, device="cuda")
    t_tbo = t_tbo_cpu = t_q = t_q_cpu = 0.0
    gen = torch.Generator(device="cuda")
    gen.manual_seed(99)
    posn = torch.arange(L, device="cuda", dtype=torch.int32)

    def sliced(fn, n_units, per):
        """run fn(a, b) over `cores` slices of the units (the C oracles release the GIL)"""
        edges = [per * (n_units * i // cores) for i in range(cores + 1)]
        return list(pool.map(lambda ab: fn(*ab), zip(edges[:-1], edges[1:])))

    tot = {}
    crc = 0
    done = 0
    mism = 0
    t_gpu = t_cpu = 0.0
    while done < args.pairs:
        m = min(cp, args.pairs - done)
        nr = 2 * m
        if args.workload in ("cfg2", "cfg4"):
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
        want, st = ora.process(hb, ho, paired, threads=cores, want_mask=want_mask)
        t_cpu += time.perf_counter() - t0
        for name in ("id0", "lo", "hi", "count", "flags"):
            got = outs[name][:nr].cpu().numpy()
            mism += int(np.count_nonzero(got != want.fields()[name]))
            crc = zlib.crc32(got.tobytes(), crc)
        if want_mask:
            got = outs["maskbits"][: nr * words].cpu().numpy().view(np.uint32)
            mism += int(np.count_nonzero(got != want.maskbits))
            crc = zlib.crc32(got.tobytes(), crc)
        for k_, v in st.as_dict().items():
            tot[k_] = tot.get(k_, 0) + v
        if args.tbo and args.workload == "cfg2":
            from oracle import tbo as otbo
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gpu.tbo_device(d_bases, None, d_off, nr, L, outs["lo"], outs["hi"], outs["flags"], None, d_tst)
            torch.cuda.synchronize()
            t_tbo += time.perf_counter() - t0
            lo_h, fl_h = want.lo, want.flags
            whi = want.hi.copy()
            t0 = time.perf_counter()

            def tbo_slice(a, b):
                hi2, _, _, st2 = otbo.process(hb, None, ho[a:b + 1], lo_h[a:b], whi[a:b], fl_h[a:b])
                whi[a:b] = hi2
                return st2
            for st2 in sliced(tbo_slice, m, 2):
                tbo_tot += st2
            t_tbo_cpu += time.perf_counter() - t0
            got_hi = outs["hi"][:nr].cpu().numpy()
            got_fl = outs["flags"][:nr].cpu().numpy()
            mism += int(np.count_nonzero(got_hi != whi)) + int(np.count_nonzero(((got_fl & 0x20) != 0) != (whi != want.hi)))
            crc = zlib.crc32(got_hi.tobytes(), crc)
            want.hi[:] = whi
            want.flags[:] = got_fl
        if args.qtrim and args.workload == "cfg2":
            from oracle import qtrim as oq
            slope = torch.randint(0, 45, (nr, 1), device="cuda", dtype=torch.int32, generator=gen)
            noise = torch.randint(-3, 4, (nr, L), device="cuda", dtype=torch.int32, generator=gen)
            d_quals = (torch.clamp(40 - (posn * slope) // L + noise, 2, 41) + 33).to(torch.uint8).reshape(-1)
            del slope, noise
            qcfg = gpu.qtrim_cfg(qtrim_left=1, qtrim_right=1, trimq=10.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gpu.qtrim_device(d_bases, d_quals, d_off, nr, True, outs["lo"], outs["hi"], outs["flags"], qcfg, d_qst)
            torch.cuda.synchronize()
            t_q += time.perf_counter() - t0
            hq = d_quals.cpu().numpy()
            qp = oq.params(qtrim="rl", trimq=10.0)
            wl, wh, wf = want.lo.copy(), want.hi.copy(), want.flags.copy()
            t0 = time.perf_counter()

            def q_slice(a, b):
                l2, h2, f2, st2 = oq.process(hb, hq, ho[a:b + 1], True, wl[a:b], wh[a:b], wf[a:b], qp)
                wl[a:b], wh[a:b], wf[a:b] = l2, h2, f2
                return st2
            for st2 in sliced(q_slice, m, 2):
                q_tot += st2
            t_q_cpu += time.perf_counter() - t0
            for name, w in (("lo", wl), ("hi", wh), ("flags", wf)):
                got = outs[name][:nr].cpu().numpy()
                mism += int(np.count_nonzero(got != w))
                crc = zlib.crc32(got.tobytes(), crc)
            del d_quals
        done += m
    dev_tot = dict(zip(tot.keys(), d_stats.cpu().tolist()))
    ro, bo = ora.scaffold_counts()
    rg, bg = gpu.scaffold_counts()
    extra = {}
    if args.tbo:
        extra["tbo"] = {"reads_trimmed": int(tbo_tot[0]), "bases_trimmed": int(tbo_tot[1]), "counters_equal": d_tst.cpu().tolist() == tbo_tot.tolist(),
                        "gpu_s": round(t_tbo, 3), "oracle_s": round(t_tbo_cpu, 3)}
        assert extra["tbo"]["counters_equal"]
    if args.qtrim:
        extra["qtrim"] = {"stats8": q_tot.tolist(), "counters_equal": d_qst.cpu().tolist() == q_tot.tolist(), "gpu_s": round(t_q, 3),
                          "oracle_s": round(t_q_cpu, 3), "command": "qtrim=rl trimq=10, synthetic qualities decaying from Q40 (torch generator seed 99)"}
        assert extra["qtrim"]["counters_equal"]
    if args.workload == "cfg2" and args.mode != "ktrimr":
        extra["mode"] = args.mode
        extra["engine_flags"] = kw
    print(json.dumps({"workload": args.workload, "reads": 2 * args.pairs, "stored_kmers": stored, "mismatching_fields": mism, **extra,
                      "counters_equal": dev_tot == tot, "scaffold_counts_equal": bool(np.array_equal(ro, rg) and np.array_equal(bo, bg)),
                      "counters": tot, "crc32_of_results": crc, "gpu_s": round(t_gpu, 3), "oracle_s": round(t_cpu, 3),
                      "oracle_threads": cores}))
    assert mism == 0 and dev_tot == tot


if __name__ == "__main__":
    main()
