"""GPU parity of the low-entropy read filter (bbduk_b200_entropy / _entropy_device) against the oracle: flags, kept
lengths (trimfailuresto1bp) and the two counters, bit for bit -- the decision hangs on a double-precision running sum
and two float casts, all in the reference's order."""
import numpy as np
import pytest

from bbtools_b200 import make_cfg
from bbtools_b200._abi import Outputs
from oracle import entropy as oe
from test_entropy_oracle import entropy_batch

pytestmark = pytest.mark.gpu


def engine(rieb=True, tf1=False, **_):
    from bbtools_b200.bbduk import BBDukIndexGPU
    return BBDukIndexGPU(make_cfg(k=23, ktrim_right=1, require_both_bad=int(not rieb), trim_failures_to_1bp=int(tf1)))


CASES = [dict(cutoff=0.5), dict(cutoff=0.7, rieb=False), dict(cutoff=0.3, tf1=True), dict(cutoff=0.9, k=4, window=30),
         dict(cutoff=0.5, high_pass=False), dict(cutoff=-1.0), dict(cutoff=0.6, k=2, window=12), dict(cutoff=0.8, k=5, window=258)]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_ragged_reads(case):
    c = CASES[case]
    paired = case % 2 == 0
    bases, offsets, lo, hi, flags = entropy_batch(8000, 300 + case, L=160)
    if paired:
        flags[0::2][np.arange(4000) % 23 == 0] = 2
        flags[1::2][np.arange(4000) % 23 == 0] = 2
    else:
        flags[np.arange(8000) % 31 == 0] = 2
    g = engine(**c)
    p = oe.params(**c)
    whi, wfl, wst = oe.process(bases, offsets, paired, lo, hi, flags, p)
    out = Outputs(len(lo))
    out.lo[:], out.hi[:], out.flags[:] = lo, hi, flags
    st = g.entropy(bases, offsets, paired, out, g.entropy_cfg(cutoff=c["cutoff"], k=c.get("k", 5), window=c.get("window", 50),
                                                              high_pass=int(c.get("high_pass", True))))
    assert np.array_equal(out.flags, wfl), f"{np.count_nonzero(out.flags != wfl)} flags differ"
    assert np.array_equal(out.hi, whi) and list(st) == list(wst)
    if c["cutoff"] > 0 and c.get("high_pass", True):
        assert wst[0] > 200


def test_device_entry_point_and_limits():
    import torch
    g = engine()
    bases, offsets, lo, hi, flags = entropy_batch(20000, 9, L=150)
    whi, wfl, wst = oe.process(bases, offsets, True, lo, hi, flags, oe.params(cutoff=0.55))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_hi, d_fl = d(hi), d(flags)
    d_st = torch.zeros(2, dtype=torch.int64, device="cuda")
    g.entropy_device(d(bases), d(offsets.astype(np.int32)), len(lo), True, d(lo), d_hi, d_fl, g.entropy_cfg(cutoff=0.55), d_st)
    torch.cuda.synchronize()
    assert np.array_equal(d_fl.cpu().numpy(), wfl) and np.array_equal(d_hi.cpu().numpy(), whi) and d_st.cpu().tolist() == list(wst)
    out = Outputs(len(lo))
    out.lo[:], out.hi[:], out.flags[:] = lo, hi, flags
    with pytest.raises(RuntimeError, match="no device path"):
        g.entropy(bases, offsets, True, out, g.entropy_cfg(cutoff=0.5, k=6))


MASK_CASES = [dict(cutoff=0.5), dict(cutoff=0.75, k=4, window=30, tf1=True), dict(cutoff=0.35, k=3, window=20), dict(cutoff=0.6, high_pass=False, k=2, window=12),
              dict(cutoff=0.8, k=5, window=120)]


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("case", range(len(MASK_CASES)))
def test_entropy_mask_and_trim(case, mode):
    """bbduk_b200_entropy_mask: mark bits, lo / hi (trim mode) and the two counters against the oracle, bit for bit"""
    c = MASK_CASES[case]
    paired = case % 2 == 0
    bases, offsets, lo, hi, flags = entropy_batch(8000, 500 + case, L=170)
    if paired:
        flags[0::2][np.arange(4000) % 19 == 0] = 2
        flags[1::2][np.arange(4000) % 19 == 0] = 2
    g = engine(**c)
    wlo, whi, wbits, woff, wst = oe.mask(bases, offsets, paired, lo, hi, flags, oe.params(**c), mode)
    out = Outputs(len(lo))
    out.lo[:], out.hi[:], out.flags[:] = lo, hi, flags
    bits, moff, st = g.entropy_mask(bases, offsets, paired, out, g.entropy_cfg(cutoff=c["cutoff"], k=c.get("k", 5), window=c.get("window", 50),
                                                                              high_pass=int(c.get("high_pass", True))), mode)
    assert np.array_equal(moff, woff)
    assert np.array_equal(bits, wbits), f"{np.count_nonzero(bits != wbits)} mask words differ"
    assert np.array_equal(out.lo, wlo) and np.array_equal(out.hi, whi) and np.array_equal(out.flags, flags)
    assert list(st) == list(wst) and wst[1] > 1000
