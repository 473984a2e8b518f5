#!/usr/bin/env python
"""Randomized GPU-vs-oracle stress of the k-mer block beyond the sizes of tests/test_parity_gpu.py: ragged reads (0..300
bases) with adapter fragments of either strand at random places, alphabets from plain A C G T to N-heavy / IUPAC / lower
case, paired and single, through the modes the tuned kernel serves (ktrim r / l, kfilter, ktrim=N) with hdist 0 / 1, short
k-mers on / off, trim padding, exclusive trimming -- every output field, the counters and the scaffold counts.
    python tools/stress_kmer_gpu.py FIRST_SEED LAST_SEED [reads per case]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = [
    dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1),
    dict(k=23, mink=11, hdist=1, ktrim_left=1),
    dict(k=23, hdist=1),
    dict(k=23, mink=11, hdist=1, ktrim_n=1),
    dict(k=23, ktrim_right=1),
    dict(k=21, hdist=0, ktrim_right=1, forbid_ns=1, mask_middle=0, min_len_fraction=0.5),
    dict(k=25, mink=9, hdist=1, hdist2=0, ktrim_right=1, ktrim_exclusive=1, trim_pad=2),
    dict(k=27, mink=12, hdist=1, ktrim_left=1, trim_pad=1),
    dict(k=19, hdist=0, mask_middle=0),
    dict(k=31, mink=15, hdist=0, ktrim_n=1, mask_middle=0),
]
ALPHABETS = [b"ACGT", b"ACGTN", b"ACGTNNNN", b"ACGTNacgtnRYKMUu", b"ACGTacgt"]


def adapter_concatenations(rng, rb, roff, n_reads, alpha):
    """reads made of adapter sequences of either strand glued together (every window a candidate, most of them hits), with
    a few point changes from `alpha`: the worst case for the candidate queues of the tuned kernel"""
    comp = np.arange(256, dtype=np.uint8)
    for x, y in zip(b"ACGTacgt", b"TGCAtgca"):
        comp[x] = y
    al = np.frombuffer(alpha, np.uint8)
    n_ref = len(roff) - 1
    seqs = []
    for _ in range(n_reads):
        want = int(rng.integers(0, 301))
        parts, have = [], 0
        while have < want:
            a = int(rng.integers(0, n_ref))
            frag = rb[roff[a]:roff[a + 1]]
            if rng.random() < 0.5:
                frag = comp[frag[::-1]]
            parts.append(frag)
            have += len(frag)
        s = (np.concatenate(parts)[:want] if parts else np.zeros(0, np.uint8)).copy()
        hit = rng.random(len(s)) < rng.choice([0.0, 0.01, 0.04])
        s[hit] = al[rng.integers(0, len(al), int(hit.sum()))]
        seqs.append(s)
    off = np.zeros(n_reads + 1, np.int64)
    np.cumsum([len(x) for x in seqs], out=off[1:])
    return (np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)).astype(np.uint8), off


def main():
    from test_parity_gpu import assert_same, check_scaffold_counts

    from bbtools_b200 import make_cfg, synth
    from bbtools_b200.bbduk import BBDukIndexGPU
    from bbtools_b200.fasta import read_fasta
    from oracle.oracle import Oracle
    first, last = int(sys.argv[1]), int(sys.argv[2])
    n_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 60000
    names, rb, roff = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
    n_ref = len(roff) - 1
    cases = 0
    for seed in range(first, last):
        rng = np.random.default_rng(seed)
        for ci, kw in enumerate(CONFIGS):
            cfg = make_cfg(**kw)
            o, g = Oracle(cfg), BBDukIndexGPU(cfg)
            o.add_ref(rb, roff)
            g.add_ref(rb, roff)
            assert o.finalize() == g.finalize()
            alpha = ALPHABETS[int(rng.integers(0, len(ALPHABETS)))]
            a = int(rng.integers(0, n_ref))
            adapter = rb[roff[a]:roff[a + 1]]
            if (seed + ci) % 3 == 2:
                b, off = adapter_concatenations(rng, rb, roff, n_reads // 4, alpha)
            else:
                b, off = synth.ragged_reads(n_reads, seed=1000 * seed + ci, min_len=0, max_len=300, alphabet=alpha, adapter=adapter)
            paired = bool(rng.integers(0, 2))
            want_mask = bool(kw.get("ktrim_n"))
            assert_same(o, g, b, off, paired, want_mask=want_mask, threads=16)
            check_scaffold_counts(o, g)
            g.close()
            cases += 1
        print(f"seed {seed} ok ({cases} cases of {n_reads} reads)", flush=True)


if __name__ == "__main__":
    main()
