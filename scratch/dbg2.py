import sys, numpy as np
sys.path.insert(0, '.')
from bbtools_b200 import make_cfg, synth
from bbtools_b200.bbduk import BBDukIndexGPU
from bbtools_b200.fasta import read_fasta
from oracle.oracle import Oracle
_, rb, roff = read_fasta("tests/golden/adapters.fa")
cfg = make_cfg(k=23, mink=11, hdist=1, ktrim_left=1)
g = BBDukIndexGPU(cfg); g.add_ref(rb, roff); g.finalize()
o = Oracle(cfg); o.add_ref(rb, roff); o.finalize()
b, off = synth.paired_adapter_reads(6000, seed=3)
b, off = b[:off[64]], off[:65]
eg, sg = g.process(b, off, True)
eo, so = o.process(b, off, True)
print("lo", eo.lo[20:32], eg.lo[20:32])
print(bytes(b[off[26]:off[27]]).decode())
