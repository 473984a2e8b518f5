#!/usr/bin/env python
"""Time the k-mer block on device-resident synthetic batches (CUDA events), for A/B runs of kernel changes:

    python tools/time_kmer_block.py [--pairs N] [--kinds cfg2,noN,random,polyA] [--cfg k=23,mink=11,hdist=1,ktrim_right=1,trim_pairs_evenly=1]

BBDUK_B200_FAST2=0 selects the round-1 kernel. Prints one JSON object; with --check compares a slice with the oracle."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4 << 20)
    ap.add_argument("--kinds", default="cfg2,noN,random,polyA")
    ap.add_argument("--cfg", default="k=23,mink=11,hdist=1,ktrim_right=1,trim_pairs_evenly=1")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--check", type=int, default=20000, help="pairs compared with the oracle (0 = none)")
    ap.add_argument("--counters", action="store_true", help="print the debug counters of a -DBB_FAST_COUNT build")
    a = ap.parse_args()
    import torch
    from bbtools_b200 import _lib, make_cfg
    from bbtools_b200.bbduk import BBDukIndexGPU
    from bbtools_b200.fasta import read_fasta
    lib = _lib.load()
    kw = {}
    for item in a.cfg.split(","):
        k_, v = item.split("=")
        kw[k_] = float(v) if "." in v else int(v)
    cfg = make_cfg(**kw)
    _, rb, roff = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
    eng = BBDukIndexGPU(cfg)
    eng.add_ref(rb, roff)
    stored = eng.finalize()
    L = 150
    eng.set_max_read_len(L)
    n_pairs, n_reads = a.pairs, 2 * a.pairs
    dev = torch.device("cuda", 0)
    res = {"stored": stored, "fast2": os.environ.get("BBDUK_B200_FAST2", "1"), "pairs": n_pairs}
    outs = {"id0": torch.empty(n_reads, dtype=torch.int32, device=dev), "hi": torch.empty(n_reads, dtype=torch.int32, device=dev),
            "flags": torch.empty(n_reads, dtype=torch.uint8, device=dev)}
    d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
    for kind in a.kinds.split(","):
        d_bases = torch.empty(n_reads * L, dtype=torch.uint8, device=dev)
        d_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
        if kind == "cfg2":
            assert lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), n_pairs, 0, L, C.c_uint64(1), 50, 5, None) == 0
        elif kind == "noN":
            assert lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), n_pairs, 0, L, C.c_uint64(1), 50, 0, None) == 0
        elif kind == "random":
            g = torch.Generator(device=dev)
            g.manual_seed(7)
            codes = torch.randint(0, 4, (n_reads * L,), device=dev, dtype=torch.uint8, generator=g)
            d_bases.copy_(torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[codes.long()])
            d_off.copy_(torch.arange(0, (n_reads + 1) * L, L, dtype=torch.int32, device=dev))
            del codes
        elif kind == "polyA":
            d_bases.fill_(65)
            d_off.copy_(torch.arange(0, (n_reads + 1) * L, L, dtype=torch.int32, device=dev))
        else:
            raise SystemExit("unknown kind " + kind)
        torch.cuda.synchronize()
        for _ in range(3):
            eng.process_device(d_bases, d_off, n_reads, True, outs, d_stats=d_stats)
        torch.cuda.synchronize()
        if a.counters:
            z = (C.c_ulonglong * 16)()
            for name in ("bbduk_b200_debug_fast2_counters", "bbduk_b200_debug_fast_counters"):
                if hasattr(lib, name):
                    getattr(lib, name)(z, 1)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
        ev[0].record()
        for i in range(a.reps):
            eng.process_device(d_bases, d_off, n_reads, True, outs, d_stats=d_stats)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))
        res[kind] = {"ms_median": ms[len(ms) // 2], "ms_min": ms[0], "greads_per_s": n_reads / ms[len(ms) // 2] / 1e6}
        if a.counters:
            for name in ("bbduk_b200_debug_fast2_counters", "bbduk_b200_debug_fast_counters"):
                if hasattr(lib, name):
                    getattr(lib, name)(z, 1)
                    res[kind][name] = [int(x) / (a.reps * n_reads / 32) for x in z]
        if a.check and kind in ("cfg2", "noN", "random"):
            from oracle.oracle import Oracle
            o = Oracle(cfg)
            o.add_ref(rb, roff)
            o.finalize()
            chk = min(a.check, n_pairs)
            hb = d_bases[: 2 * chk * L].cpu().numpy()
            ho = np.arange(0, (2 * chk + 1) * L, L, dtype=np.int64)
            want, _ = o.process(hb, ho, True, threads=8)
            bad = {}
            for name in ("hi", "id0", "flags"):
                got = outs[name][: 2 * chk].cpu().numpy()
                w = getattr(want, name)
                nb = int((got != w).sum())
                if nb:
                    i = int(np.nonzero(got != w)[0][0])
                    bad[name] = {"n": nb, "first": i, "oracle": int(w[i]), "gpu": int(got[i]),
                                 "read": bytes(hb[i * L:(i + 1) * L]).decode()}
            res[kind]["mismatch"] = bad
        del d_bases, d_off
    print(json.dumps(res))


if __name__ == "__main__":
    main()
