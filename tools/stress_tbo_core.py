#!/usr/bin/env python
"""Stress of the host build of bbtools_b200/csrc/tbo_core.cuh (tests/tbo_core_host.cpp) against the tbo oracle on CPU:
tandem repeats, internal duplications, mismatch rates straddling maxRatio, N, and (odd seeds) random maxRatio / margin /
offset / minSecondRatio.  python tools/stress_tbo_core.py FIRST_SEED LAST_SEED [gpu]   (4000 pairs x 2 parameter sets per seed;
"gpu": the same pairs through bbduk_b200_tbo on the device, strict and loose defaults)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import test_tbo_core_cpu as t
from oracle import tbo as otbo
host = t.host.__wrapped__()
COMP = t.COMP
acgt = np.frombuffer(b"ACGT", np.uint8)
def gen(seed, npairs):
    rng = np.random.default_rng(seed)
    seqs = []
    for i in range(npairs):
        n1, n2 = (int(x) for x in rng.integers(30, 320, 2))
        ins = int(rng.integers(10, n1 + n2))
        kind = i % 5
        if kind == 0:
            unit = acgt[rng.integers(0, 4, int(rng.integers(1, 12)))]
            frag = np.tile(unit, ins // len(unit) + 1)[:ins].copy()
        elif kind == 1:  # fragment with an internal duplication -> two competing alignments
            half = acgt[rng.integers(0, 4, max(ins // 2, 1))]
            frag = np.concatenate([half, half])[:ins].copy()
            if len(frag) < ins: frag = np.concatenate([frag, acgt[rng.integers(0,4,ins-len(frag))]])
        else:
            frag = acgt[rng.integers(0, 4, ins)]
        if i % 7 == 3:  # the same fragment positions are N in both mates
            frag = frag.copy()
            frag[rng.random(len(frag)) < 0.05 + 0.3 * rng.random()] = ord('N')
        r1 = np.concatenate([frag, acgt[rng.integers(0, 4, 400)]])[:n1].copy()
        r2 = np.concatenate([COMP[frag[::-1]], acgt[rng.integers(0, 4, 400)]])[:n2].copy()
        rate = rng.random() * 0.25
        for r in (r1, r2):
            hit = rng.random(len(r)) < rate / 2
            r[hit] = acgt[rng.integers(0, 4, int(hit.sum()))]
            if rng.random() < 0.1: r[int(rng.integers(0, len(r)))] = ord('N')
        seqs += [r1, r2]
    bases = np.concatenate(seqs).astype(np.uint8)
    offsets = np.zeros(len(seqs) + 1, np.int64); np.cumsum([len(x) for x in seqs], out=offsets[1:])
    return bases, offsets
tot = 0
GPU = len(sys.argv) > 3 and sys.argv[3] == "gpu"  # third argument "gpu": the device kernels (default parameters only) instead of the host build
if GPU:
    import test_tbo_gpu as tg
    eng = tg.engine()
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    bases, offsets = gen(seed, 4000)
    if GPU:
        L = np.diff(offsets).astype(np.int32)
        for strict in (True, False):
            st = tg.check(eng, bases, None, offsets, np.zeros(len(L), np.int32), L, np.zeros(len(L), np.uint8), strict)
            tot += st[0]
        print(seed, "ok (gpu)", tot, flush=True)
        continue
    L = np.diff(offsets).astype(np.int32); z = np.zeros(len(L), np.int32); f = np.zeros(len(L), np.uint8)
    rng = np.random.default_rng(1000 + seed)
    for strict in (True, False):
        p = otbo.default_params(strict)
        if seed % 2:
            p.max_ratio = float(rng.choice([0.03, 0.05, 0.1, 0.15, 0.2])); p.ratio_margin = float(rng.choice([1.5, 2.0, 5.0, 9.0]))
            p.ratio_offset = float(rng.choice([0.0, 0.4, 0.5, 0.9])); p.min_second_ratio = float(rng.choice([0.05, 0.12, 0.3]))
        want = otbo.process(bases, None, offsets, z, L, f, p)
        got = t.run_host(host, bases, offsets, z, L, f, p)
        for k, (g, w) in enumerate(zip(got[:4], want)):
            assert np.array_equal(g, w), (seed, strict, k, np.flatnonzero(np.asarray(g) != np.asarray(w))[:5])
        tot += want[3][0]
    print(seed, "ok", tot, flush=True)
