#!/usr/bin/env python
"""Device-resident timing (CUDA events) of the k-mer block for one mode configuration after the other on the cfg-2 batch:
reads/s of the tuned kernels and of the generic kernel (the modes only it serves), one JSON line per configuration.

    python tools/time_modes.py [--pairs N]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MODES = [
    ("ktrim=r k=23 mink=11 hdist=1 tpe (cfg 2, tuned)", dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)),
    ("ktrim=l k=23 mink=11 hdist=1 (tuned)", dict(k=23, mink=11, hdist=1, ktrim_left=1)),
    ("k=31 kfilter (tuned)", dict(k=31)),
    ("ktrim=r k=23 (cfg 1 flags: hdist 0, maskmiddle, forbidNs; tuned)", dict(k=23, ktrim_right=1)),
    ("ktrim tips k=23 mink=11 hdist=1 (generic)", dict(k=23, mink=11, hdist=1, ktrim_left=1, ktrim_right=1)),
    ("kmask ktrim=N k=23 mink=11 hdist=1 (tuned since r02l)", dict(k=23, mink=11, hdist=1, ktrim_n=1)),
    ("kfilter mbk=2 k=31 (generic)", dict(k=31, max_bad_kmers=2, mask_middle=0)),
    ("kfilter mcf=0.2 k=25 (generic)", dict(k=25, min_covered_fraction=0.2)),
    ("findbestmatch k=25 (generic)", dict(k=25, find_best_match=1)),
    ("k=40 countSetKmersBig (generic)", dict(k=40)),
    ("qhdist=1 k=23 ktrim=r (generic)", dict(k=23, qhdist=1, ktrim_right=1)),
    ("speed=5 k=20 (generic)", dict(k=20, speed=5)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from bbtools_b200 import _lib, make_cfg
    from bbtools_b200.bbduk import BBDukIndexGPU
    from bbtools_b200.fasta import read_fasta
    lib = _lib.load()
    _, rb, roff = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
    L, n_pairs = 150, a.pairs
    n_reads = 2 * n_pairs
    dev = torch.device("cuda", 0)
    d_bases = torch.empty(n_reads * L, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
    assert lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), n_pairs, 0, L, C.c_uint64(1), 50, 5, None) == 0
    for name, kw in MODES:
        eng = BBDukIndexGPU(make_cfg(**kw))
        eng.add_ref(rb, roff)
        stored = eng.finalize()
        eng.set_max_read_len(L)
        outs = {"id0": torch.empty(n_reads, dtype=torch.int32, device=dev), "lo": torch.empty(n_reads, dtype=torch.int32, device=dev),
                "hi": torch.empty(n_reads, dtype=torch.int32, device=dev), "count": torch.empty(n_reads, dtype=torch.int32, device=dev),
                "flags": torch.empty(n_reads, dtype=torch.uint8, device=dev)}
        paired = not kw.get("ksplit")
        if kw.get("ktrim_n"):
            words = (L + 31) // 32
            outs["mask_off"] = torch.arange(0, (n_reads + 1) * words, words, dtype=torch.int64, device=dev)
            outs["maskbits"] = torch.zeros(n_reads * words, dtype=torch.int32, device=dev)
        d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
        eng.process_device(d_bases, d_off, n_reads, paired, outs, d_stats=d_stats)
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.process_device(d_bases, d_off, n_reads, paired, outs, d_stats=d_stats)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(json.dumps({"mode": name, "stored_kmers": stored, "reads": n_reads, "ms": round(ms, 3), "reads_per_s": round(n_reads / ms * 1e3)}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
