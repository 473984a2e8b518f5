#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (tools/run_sanitizers.sh): table build (atomicCAS / atomicMin
inserts, neighbourhood expansion), the tuned per-read kernels (shared-memory atomics, pooled rounds), the generic kernel, the
HBM-table direct path, tbo / qtrim / entropy, and the counting table with a resize. Every result is checked against the oracle,
so a sanitizer run is also a parity run. Sizes are tiny: the tools slow kernels down 10-1000x."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bbtools_b200 import make_cfg, synth  # noqa: E402
from bbtools_b200.bbduk import BBDukIndexGPU  # noqa: E402
from bbtools_b200.fasta import read_fasta  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def same(o, g, b, off, paired, mask=False):
    eo, so = o.process(b, off, paired, threads=4, want_mask=mask)
    eg, sg = g.process(b, off, paired, want_mask=mask)
    for name, x in eo.fields().items():
        assert np.array_equal(x, eg.fields()[name]), name
    assert so.as_dict() == sg.as_dict()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    _, rb, roff = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
    n = int(os.environ.get("SAN_PAIRS", "1500"))
    b, off = synth.paired_adapter_reads(n, seed=1)
    rgb, rgo = synth.ragged_reads(800, seed=4, adapter=b"AGATCGGAAGAGCACACGTCTGAACTCCAGTCA")
    if which in ("all", "bbduk"):
        for kw in (dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1),  # table build hdist 1 + tails, fast2 ktrim r
                   dict(k=23, mink=11, hdist=1, ktrim_left=1),                       # fast2 ktrim l (downward rounds)
                   dict(k=31),                                                       # fast2 kfilter, maskmiddle
                   dict(k=21, hdist=0, ktrim_right=1, forbid_ns=1, mask_middle=0),   # forbidNs: forced windows
                   dict(k=23, mink=11, hdist=1, ktrim_left=1, ktrim_right=1),        # generic kernel: tips
                   dict(k=23, mink=11, hdist=1, ktrim_n=1),                          # generic kernel: kmask
                   dict(k=25, hdist=1, edist=1, ktrim_right=1)):                     # table build with indels
            mask = bool(kw.get("ktrim_n"))
            o, g = Oracle(make_cfg(**kw)), BBDukIndexGPU(make_cfg(**kw))
            o.add_ref(rb, roff)
            g.add_ref(rb, roff)
            assert o.finalize() == g.finalize()
            same(o, g, b, off, True, mask)
            same(o, g, rgb, rgo, False, mask)
            print("ok", kw, flush=True)
    if which in ("all", "direct"):
        ref_b, ref_off = synth.random_reference(40, 60000, seed=7)  # 2.4 M keys: too many for the on-chip filters -> direct path
        kw = dict(k=31)
        o, g = Oracle(make_cfg(**kw)), BBDukIndexGPU(make_cfg(**kw))
        o.add_ref(ref_b, ref_off)
        g.add_ref(ref_b, ref_off)
        assert o.finalize() == g.finalize()
        cb, co = synth.contaminant_reads(3000, ref_b, seed=1)
        l0 = g.launches
        same(o, g, cb, co, False)
        assert g.launches - l0 >= 5, "expected the direct path"
        print("ok direct", flush=True)
    if which in ("all", "steps"):
        from oracle import qtrim as oq
        from oracle import tbo as otbo
        kw = dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
        o, g = Oracle(make_cfg(**kw)), BBDukIndexGPU(make_cfg(**kw))
        o.add_ref(rb, roff)
        g.add_ref(rb, roff)
        o.finalize(), g.finalize()
        want, _ = o.process(b, off, True, threads=4)
        got, _ = g.process(b, off, True)
        rng = np.random.default_rng(3)
        q = (33 + np.clip(40 - (np.arange(len(b)) % 150) * rng.integers(0, 45, len(b)) // 150, 2, 41)).astype(np.uint8)
        whi, _, _, wt = otbo.process(b, q, off, want.lo, want.hi, want.flags)
        _, gt = g.tbo(b, q, off, got)
        assert np.array_equal(got.hi, whi) and list(gt) == list(wt)
        wl, wh, wf, wq = oq.process(b, q, off, True, want.lo, whi, got.flags, oq.params(qtrim="rl", trimq=10.0))
        gq = g.qtrim(b, q, off, True, got, g.qtrim_cfg(qtrim_left=1, qtrim_right=1, trimq=10.0))
        assert np.array_equal(got.lo, wl) and np.array_equal(got.hi, wh) and np.array_equal(got.flags, wf) and list(gq) == list(wq)
        from oracle import entropy as oe
        ecfg = g.entropy_cfg(cutoff=0.7)
        ge = g.entropy(b, off, True, got, ecfg)
        print("ok tbo qtrim entropy", list(gt), list(gq)[:2], list(ge), flush=True)
        # the same steps through the chain entry (packed upload + unpack_kernel when the batch is plain A C G T N)
        co, cst, ct2, cq8, ce2 = g.process_chain(b, q, off, True, tbo=g.tbo_cfg(), qtrim=g.qtrim_cfg(qtrim_left=1, qtrim_right=1, trimq=10.0),
                                                  entropy=ecfg)
        assert np.array_equal(co.lo, got.lo) and np.array_equal(co.hi, got.hi) and np.array_equal(co.flags, got.flags)
        assert list(ct2) == list(gt) and list(cq8) == list(gq) and list(ce2) == list(ge)
        print("ok chain", flush=True)
    if which in ("all", "kcount"):
        from bbtools_b200.kcount import KmerTableSetGPU
        from oracle.kcount import KCountOracle
        gb, go = synth.genome_reads(4000, 200000, seed=11)
        t = KmerTableSetGPU(31, True, initial_keys=1024)  # tiny table: several resizes (rehash kernel)
        ko = KCountOracle(31, True)
        for c in range(4):
            sl = slice(c * 1000, (c + 1) * 1000 + 1)
            bb = gb[go[sl][0]:go[sl][-1]]
            oo = go[sl] - go[sl][0]
            t.add_reads(bb, oo)
            ko.add_reads(bb, oo)
        assert t.stats()["unique_kmers"] == ko.stats()["unique_kmers"]
        kg, cg = t.dump()
        kk, ck = ko.dump()
        assert np.array_equal(np.sort(kg), np.sort(kk))
        print("ok kcount", t.stats(), flush=True)


if __name__ == "__main__":
    main()
