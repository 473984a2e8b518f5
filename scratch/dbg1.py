import sys, numpy as np
sys.path.insert(0, '.')
from bbtools_b200 import make_cfg, synth
from bbtools_b200.bbduk import BBDukIndexGPU
from bbtools_b200.fasta import read_fasta
from oracle.oracle import Oracle
_, rb, roff = read_fasta("tests/golden/adapters.fa")
for kw in (dict(k=23, ktrim_right=1), dict(k=23, mink=11, hdist=1, ktrim_right=1), dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)):
    cfg = make_cfg(**kw)
    g = BBDukIndexGPU(cfg); g.add_ref(rb, roff); g.finalize()
    o = Oracle(cfg); o.add_ref(rb, roff); o.finalize()
    for n, paired in ((64, True), (64, False), (3000, True)):
        b, off = synth.paired_adapter_reads(n, seed=3)
        eg, sg = g.process(b, off, paired)
        eo, so = o.process(b, off, paired)
        ok = all(np.array_equal(x, eg.fields()[k]) for k, x in eo.fields().items())
        print(kw, n, paired, "arrays", ok, "stats", so.as_dict() == sg.as_dict(), sg.as_dict() if so.as_dict() != sg.as_dict() else "")
