#!/usr/bin/env python
"""The host-buffer path across its chunk boundaries (512 Ki reads / 256 MiB per chunk, ASCII and packed chunks mixed): a
batch of ragged pairs (cfg-2 pairs cut to random lengths 0..150, so that reads and pairs straddle every boundary) through
bbduk_b200_process, pageable and pinned input, and through bbduk_b200_process_packed, against the 16-thread oracle.
    python tools/stress_chunks_gpu.py [pairs, default 1500000] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    from bbtools_b200 import make_cfg, synth
    from bbtools_b200.bbduk import BBDukIndexGPU
    from bbtools_b200.fasta import read_fasta
    from oracle.oracle import Oracle
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    _, rb, roff = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
    full, foff = synth.paired_adapter_reads(n_pairs, seed=seed)
    n = len(foff) - 1
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 151, n)
    lens[rng.random(n) < 0.5] = 150
    keep = (np.arange(150)[None, :] < lens[:, None]).reshape(-1)
    bases = np.ascontiguousarray(full[keep])
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    del full, keep
    for kw in (dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1), dict(k=23, hdist=1)):
        cfg = make_cfg(**kw)
        o, g = Oracle(cfg), BBDukIndexGPU(cfg)
        o.add_ref(rb, roff)
        g.add_ref(rb, roff)
        assert o.finalize() == g.finalize()
        want, wst = o.process(bases, offsets, True, threads=os.cpu_count() or 1)
        hb = torch.from_numpy(bases).pin_memory().numpy()
        ho = torch.from_numpy(offsets).pin_memory().numpy()
        F, D = g.pack(bases)
        runs = [("pageable", lambda: g.process(bases, offsets, True)), ("pinned", lambda: g.process(hb, ho, True)),
                ("pinned again (adaptive chunk mix warmed up)", lambda: g.process(hb, ho, True)),
                ("packed by the caller", lambda: g.process_packed(F, D, offsets, True))]
        for name, fn in runs:
            got, gst = fn()
            for field, x in want.fields().items():
                y = got.fields()[field]
                assert np.array_equal(x, y), (kw, name, field, int(np.count_nonzero(x != y)))
            assert wst.as_dict() == gst.as_dict(), (kw, name)
            print("ok", kw, name, gst.as_dict()["reads_ktrimmed"], gst.as_dict()["reads_kfiltered"], flush=True)
        g.close()


if __name__ == "__main__":
    main()
