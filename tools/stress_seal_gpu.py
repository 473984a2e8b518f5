"""Randomized Seal parity beyond the fixed option sets of tests/test_seal_gpu.py: random flag combinations, references that
share stretches, ragged reads with N / IUPAC / lower case / chimeras, paired and unpaired, GPU (C ABI) vs the oracle --
table, per-unit outputs, totals and per-reference counters, bit for bit.

  python tools/stress_seal_gpu.py [--cases 40] [--seed 1]        prints one JSON line"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bbtools_b200 import seal as PS  # noqa: E402
from oracle import seal as S  # noqa: E402
from test_seal_oracle import make_case, pack  # noqa: E402


def random_flags(rng):
    k = int(rng.choice([5, 11, 13, 16, 17, 21, 24, 27, 31]))
    kw = dict(k=k)
    kw["hdist"] = int(rng.choice([0, 0, 1, 2])) if k <= 13 else int(rng.choice([0, 0, 0, 1]))
    mm = int(rng.integers(0, 3))
    if mm == 0 or k < 5:
        kw["mask_middle"] = 0
    elif mm == 2:
        kw["mid_mask_len"] = int(rng.integers(1, max(2, min(6, k - 2))))
    kw["rcomp"] = int(rng.integers(0, 4) > 0)
    kw["forbid_ns"] = int(rng.integers(0, 2))
    kw["ambig_mode"] = int(rng.integers(1, 5))
    kw["match_mode"] = int(rng.choice([1, 1, 1, 2, 3]))
    kw["keep_pairs_together"] = int(rng.integers(0, 2))
    if rng.integers(0, 2):
        kw["clearzone"] = int(rng.integers(0, 12))
    if rng.integers(0, 3) == 0:
        kw["clearzone_fraction"] = float(rng.choice([0.01, 0.05, 0.2]))
    if rng.integers(0, 3) == 0:
        kw["min_kmer_hits"] = int(rng.integers(1, 8))
    if rng.integers(0, 3) == 0:
        kw["min_kmer_fraction"] = float(rng.choice([0.05, 0.2, 0.5]))
    if rng.integers(0, 4) == 0:
        kw["restrict_left"] = int(rng.integers(20, 120))
    if rng.integers(0, 4) == 0:
        kw["restrict_right"] = int(rng.integers(20, 120))
    extra = int(rng.integers(0, 6))
    if extra == 0:
        kw["qskip"] = int(rng.integers(2, 5))
    elif extra == 1:
        kw["speed"] = int(rng.integers(1, 12))
    elif extra == 2:
        kw["rskip"] = int(rng.integers(2, 6))
    kw["ids_stride"] = int(rng.choice([0, 1, 4, 9]))
    return kw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    bad = []
    units = 0
    for c in range(a.cases):
        kw = random_flags(rng)
        cfg = PS.make_cfg(**kw)
        small = cfg.hdist == 2
        paired = bool(rng.integers(0, 2))
        refs, reads = make_case(int(rng.integers(1, 1 << 30)), n_refs=4 if small else int(rng.integers(2, 12)),
                                ref_len=150 if small else int(rng.integers(100, 600)), n_frag=int(rng.integers(50, 500)),
                                read_len=int(rng.integers(40, 400)), paired=paired, k=cfg.k, n_rate=float(rng.choice([0.0, 0.01, 0.05])))
        g, o = PS.SealIndexGPU(cfg), S.SealOracle(cfg)
        rb, ro = pack(refs)
        g.add_ref(rb, ro)
        o.add_ref(rb, ro)
        ok = g.finalize() == o.finalize()
        gk, gi = g.table()
        okk, oi = o.table()
        ok = ok and np.array_equal(gk, okk) and np.array_equal(gi, oi)
        b, off = pack(reads)
        first = int(rng.integers(0, 1 << 40))
        gr, gs = g.process(b, off, paired, first)
        wr, ws = o.process(b, off, paired, first)
        ok = ok and all(np.array_equal(x, gr.fields()[n]) for n, x in wr.fields().items()) and gs.as_dict() == ws.as_dict()
        ok = ok and all(np.array_equal(x, y) for x, y in zip(g.scaffold_counts(), o.scaffold_counts()))
        units += len(wr.n_assigned)
        if not ok:
            bad.append({"case": c, "flags": kw, "paired": paired})
        g.close()
    print(json.dumps({"tool": "stress_seal_gpu", "seed": a.seed, "cases": a.cases, "units": units, "mismatching_cases": bad}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
