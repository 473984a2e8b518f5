"""bbduk_b200_process_chain on a call of several chunks: chunks made of A C G T N only cross PCIe 2-bit packed and are spelled
out again on the device, the first chunk with anything else (here lower case) sends the rest of the call as ASCII -- the
results must be those of the oracles either way, and the wire bytes must show both kinds of chunk."""
import numpy as np
import pytest

from bbtools_b200 import make_cfg, synth

pytestmark = pytest.mark.gpu


def test_chain_packed_and_ascii_chunks(adapters):
    from bbtools_b200.bbduk import BBDukIndexGPU
    from oracle import entropy as oe
    from oracle import qtrim as oq
    from oracle.oracle import Oracle
    _, rb, roff = adapters
    kw = dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
    o, g = Oracle(make_cfg(**kw)), BBDukIndexGPU(make_cfg(**kw))
    o.add_ref(rb, roff)
    g.add_ref(rb, roff)
    assert o.finalize() == g.finalize()
    n_pairs = 450_000  # 900 k reads = 4 chunks of 256 Ki reads
    bases, offsets = synth.paired_adapter_reads(n_pairs, seed=5)
    bases = bases.copy()
    rng = np.random.default_rng(1)
    first_odd = 600_000 * 150  # inside the third chunk
    idx = first_odd + rng.integers(0, len(bases) - first_odd, 200_000)
    bases[idx] |= 0x20  # lower case: defined, but F + D would lose the case
    quals = (33 + np.clip(40 - (np.arange(len(bases)) % 150) * rng.integers(0, 45, len(bases)) // 150, 2, 41)).astype(np.uint8)
    want, wst = o.process(bases, offsets, True, threads=8)
    wl, wh, wf, wq = oq.process(bases, quals, offsets, True, want.lo, want.hi, want.flags, oq.params(qtrim="rl", trimq=10.0))
    eh, ef, we = oe.process(bases, offsets, True, wl, wh, wf, oe.params(cutoff=0.5))
    x0 = g.transfer_bytes()
    out, st, _, q8, e2 = g.process_chain(bases, quals, offsets, True, qtrim=g.qtrim_cfg(qtrim_left=1, qtrim_right=1, trimq=10.0),
                                         entropy=g.entropy_cfg(cutoff=0.5))
    x1 = g.transfer_bytes()
    assert np.array_equal(out.lo, wl) and np.array_equal(out.hi, eh) and np.array_equal(out.flags, ef)
    assert st.as_dict() == wst.as_dict() and list(q8) == list(wq) and list(e2) == list(we)
    up = x1[0] - x0[0]
    all_ascii = 2 * len(bases) + 4 * len(offsets)
    all_packed = len(bases) + 3 * len(bases) // 8 + 4 * len(offsets)
    assert all_packed < up < all_ascii, (all_packed, up, all_ascii)
    # a second call starts packing again (the decision is per call) and is still right
    out2, st2, _, _, _ = g.process_chain(bases[: 300_000 * 150], None, offsets[: 300_001], True, entropy=g.entropy_cfg(cutoff=0.5))
    eh2, ef2, _ = oe.process(bases[: 300_000 * 150], offsets[: 300_001], True, want.lo[:300_000], want.hi[:300_000], want.flags[:300_000],
                             oe.params(cutoff=0.5))
    assert np.array_equal(out2.hi, eh2) and np.array_equal(out2.flags, ef2)
