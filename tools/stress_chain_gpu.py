#!/usr/bin/env python
"""bbduk_b200_process_chain (k-mer block -> tbo -> quality trimming -> entropy filter, one upload) on ragged pairs with
decaying random qualities and low-complexity stretches, against the four oracles run one after the other.
    python tools/stress_chain_gpu.py [pairs, default 400000] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bbtools_b200 import F_TBO, make_cfg, synth
    from bbtools_b200.bbduk import BBDukIndexGPU
    from bbtools_b200.fasta import read_fasta
    from oracle import entropy as oe
    from oracle import qtrim as oq
    from oracle import tbo as otbo
    from oracle.oracle import Oracle
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    _, rb, roff = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
    full, foff = synth.paired_adapter_reads(n_pairs, seed=seed)
    n = len(foff) - 1
    rng = np.random.default_rng(seed)
    full = full.reshape(n, 150).copy()
    # low-complexity reads: a homopolymer or a short tandem repeat over a random stretch
    for i in np.flatnonzero(rng.random(n) < 0.08):
        unit = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rng.integers(1, 4)))]
        a = int(rng.integers(0, 100))
        b = int(rng.integers(a + 20, 151))
        full[i, a:b] = np.tile(unit, 150)[: b - a]
    lens = rng.integers(0, 151, n)
    lens[rng.random(n) < 0.6] = 150
    keep = (np.arange(150)[None, :] < lens[:, None]).reshape(-1)
    bases = np.ascontiguousarray(full.reshape(-1)[keep])
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    pos = (np.arange(150)[None, :] * rng.integers(0, 45, (n, 1)) // 150).reshape(-1)[keep]
    quals = (33 + np.clip(40 - pos - rng.integers(0, 6, len(pos)), 2, 41)).astype(np.uint8)
    del full, keep, pos
    cores = os.cpu_count() or 1
    for strict, kw in ((True, dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)), (False, dict(k=23, mink=11, hdist=1, ktrim_right=1))):
        cfg = make_cfg(**kw)
        o, g = Oracle(cfg), BBDukIndexGPU(cfg)
        o.add_ref(rb, roff)
        g.add_ref(rb, roff)
        assert o.finalize() == g.finalize()
        want, wst = o.process(bases, offsets, True, threads=cores)
        whi, _, _, wt = otbo.process(bases, quals, offsets, want.lo, want.hi, want.flags, otbo.default_params(strict))
        fl = want.flags | np.where(whi != want.hi, np.uint8(F_TBO), np.uint8(0))
        wl, wh, wf, wq = oq.process(bases, quals, offsets, True, want.lo, whi, fl, oq.params(qtrim="rl", trimq=10.0))
        eh, ef, we = oe.process(bases, offsets, True, wl, wh, wf, oe.params(cutoff=0.5))
        out, st, t2, q8, e2 = g.process_chain(bases, quals, offsets, True, tbo=g.tbo_cfg(strict_overlap=int(strict)),
                                              qtrim=g.qtrim_cfg(qtrim_left=1, qtrim_right=1, trimq=10.0), entropy=g.entropy_cfg(cutoff=0.5))
        for name, x, y in (("lo", wl, out.lo), ("hi", eh, out.hi), ("flags", ef, out.flags)):
            assert np.array_equal(x, y), (kw, name, int(np.count_nonzero(x != y)))
        assert st.as_dict() == wst.as_dict() and list(t2) == list(wt) and list(q8) == list(wq) and list(e2) == list(we)
        print("ok", kw, "strict" if strict else "loose", list(t2), list(q8)[:4], list(e2), flush=True)
        g.close()


if __name__ == "__main__":
    main()
