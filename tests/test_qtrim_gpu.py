"""GPU parity of poly-X trimming, quality trimming + the quality / length / N filters (bbduk_b200_qtrim / _qtrim_device) against the
oracle: kept intervals, flags and the eight counters, bit for bit -- including every single-precision comparison of
the running score."""
import numpy as np
import pytest

from bbtools_b200 import make_cfg, synth
from bbtools_b200._abi import Outputs
from oracle import qtrim as oq
from test_qtrim_oracle import CASES, NOQUAL_CASES, qual_batch

pytestmark = pytest.mark.gpu


def engine(minlen=10, mlf=0.0, rieb=True, tf1=False, **_):
    from bbtools_b200.bbduk import BBDukIndexGPU
    return BBDukIndexGPU(make_cfg(k=23, ktrim_right=1, min_read_length=minlen, min_len_fraction=mlf, require_both_bad=int(not rieb),
                                  trim_failures_to_1bp=int(tf1)))


def dev_cfg(g, qtrim="rl", trimq=6.0, mbq=0, maxns=-1, maxlen=0, qual_offset=33, polya=0, polyg=(0, 0), fpolyg=0, polyc=(0, 0),
            fpolyc=0, maxnonpoly=1, maq=0.0, maqb=0, maxnrate=1.0, mcb=0, mbf=0.0, mode=0, window=4, goodinterval=2, **_):
    return g.qtrim_cfg(qtrim_left=int("l" in qtrim), qtrim_right=int("r" in qtrim), trimq=trimq, min_base_quality=mbq, max_ns=maxns,
                       max_read_length=maxlen, qual_offset=qual_offset, trim_poly_a=polya, trim_poly_g_left=polyg[0],
                       trim_poly_g_right=polyg[1], filter_poly_g=fpolyg, trim_poly_c_left=polyc[0], trim_poly_c_right=polyc[1],
                       filter_poly_c=fpolyc, max_non_poly=maxnonpoly, min_avg_quality=maq, min_avg_quality_bases=maqb,
                       max_n_rate=maxnrate, min_consecutive_bases=mcb, min_base_frequency=mbf, trim_mode=mode, window_length=window,
                       min_good_interval=goodinterval)


def check(case, bases, quals, offsets, paired, lo, hi, flags):
    g = engine(**case)
    want = oq.process(bases, quals, offsets, paired, lo, hi, flags, oq.params(**case))
    out = Outputs(len(lo))
    out.lo[:], out.hi[:], out.flags[:] = lo, hi, flags
    st = g.qtrim(bases, quals, offsets, paired, out, dev_cfg(g, **case))
    assert np.array_equal(out.lo, want[0]), f"{np.count_nonzero(out.lo != want[0])} left ends differ"
    assert np.array_equal(out.hi, want[1]), f"{np.count_nonzero(out.hi != want[1])} right ends differ"
    assert np.array_equal(out.flags, want[2])
    assert list(st) == list(want[3])
    return want[3]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_ragged_reads(case):
    paired = case % 2 == 0
    bases, quals, offsets, lo, hi, flags = qual_batch(6000, 200 + case, L=170, paired=paired)
    st = check(CASES[case], bases, quals, offsets, paired, lo, hi, flags)
    assert st.sum() > 100


@pytest.mark.parametrize("case", range(len(NOQUAL_CASES)))
def test_reads_without_qualities(case):
    """quals = NULL: the trimming rules fall back to N's, mbq / maq do not apply (shared/TrimRead.java:352, :418, :440, :459)"""
    paired = case % 2 == 1
    bases, _, offsets, lo, hi, flags = qual_batch(5000, 400 + case, L=150, paired=paired)
    rng = np.random.default_rng(case)
    for i in np.nonzero(rng.random(len(offsets) - 1) < 0.5)[0]:
        a, b = offsets[i], offsets[i + 1]
        bases[a:a + min(int(rng.integers(0, 6)), b - a)] = ord("N")
        bases[max(a, b - int(rng.integers(0, 6))):b] = ord("N")
    check(NOQUAL_CASES[case], bases, None, offsets, paired, lo, hi, flags)


def test_numeric_qualities_and_empty_batches():
    """qual_offset = 0 (Read.quality as the Java side holds it), zero reads, reads of length 0"""
    bases, quals, offsets, lo, hi, flags = qual_batch(3000, 77, L=120, paired=True)
    case = dict(qtrim="rl", trimq=12.0, qual_offset=0, maxns=2)
    check(case, bases, (quals - 33).astype(np.uint8), offsets, True, lo, hi, flags)
    g = engine()
    z = np.zeros(0, np.int32)
    out = Outputs(0)
    assert list(g.qtrim(np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.int64), True, out, dev_cfg(g))) == [0] * 8
    assert z.size == 0


def test_after_the_kmer_block_and_tbo_on_cfg2_pairs(adapters):
    """ktrim=r k=23 mink=11 hdist=1 tpe tbo qtrim=rl trimq=10 on 2x150 pairs whose qualities decay towards the 3' end"""
    from oracle import tbo as otbo
    from oracle.oracle import Oracle
    g = engine()
    _, rb, roff = adapters
    g.add_ref(rb, roff)
    g.finalize()
    cfgk = make_cfg(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
    from bbtools_b200.bbduk import BBDukIndexGPU
    gk = BBDukIndexGPU(cfgk)
    gk.add_ref(rb, roff)
    gk.finalize()
    o = Oracle(cfgk)
    o.add_ref(rb, roff)
    o.finalize()
    bases, offsets = synth.paired_adapter_reads(30000, seed=17)
    rng = np.random.default_rng(4)
    pos = np.arange(len(bases)) % 150
    quals = np.clip(40 - (pos * rng.integers(0, 45, len(bases))) // 150 + rng.integers(-3, 4, len(bases)), 2, 41).astype(np.uint8) + 33
    out, _ = gk.process(bases, offsets, True)
    want, _ = o.process(bases, offsets, True)
    gk.tbo(bases, quals, offsets, out)
    whi, _, _, _ = otbo.process(bases, quals, offsets, want.lo, want.hi, want.flags)
    assert np.array_equal(out.hi, whi)
    case = dict(qtrim="rl", trimq=10.0)
    wl, wh, wf, wst = oq.process(bases, quals, offsets, True, want.lo, whi, out.flags, oq.params(**case))
    st = gk.qtrim(bases, quals, offsets, True, out, dev_cfg(gk, **case))
    assert np.array_equal(out.lo, wl) and np.array_equal(out.hi, wh) and np.array_equal(out.flags, wf)
    assert list(st) == list(wst) and wst[0] > 5000


def test_device_entry_point_and_alignment_error():
    import torch
    g = engine()
    bases, quals, offsets, lo, hi, flags = qual_batch(20000, 5, L=150, paired=True)
    case = dict(qtrim="rl", trimq=8.0, mbq=2, maxns=3)
    want = oq.process(bases, quals, offsets, True, lo, hi, flags, oq.params(**case))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_lo, d_hi, d_fl = d(lo), d(hi), d(flags)
    d_st = torch.zeros(8, dtype=torch.int64, device="cuda")
    d_b, d_q = d(bases), d(quals)
    g.qtrim_device(d_b, d_q, d(offsets.astype(np.int32)), len(lo), True, d_lo, d_hi, d_fl, dev_cfg(g, **case), d_st)
    torch.cuda.synchronize()
    assert np.array_equal(d_lo.cpu().numpy(), want[0]) and np.array_equal(d_hi.cpu().numpy(), want[1])
    assert np.array_equal(d_fl.cpu().numpy(), want[2]) and d_st.cpu().tolist() == list(want[3])
    with pytest.raises(RuntimeError, match="16-byte aligned"):
        g.qtrim_device(d_b[1:], d_q[1:], d(offsets.astype(np.int32)), len(lo), True, d_lo, d_hi, d_fl, dev_cfg(g, **case), d_st)
    # no quality array: the device entry trims N's instead (reads without qualities), mbq does not apply
    want = oq.process(bases, None, offsets, True, lo, hi, flags, oq.params(**case))
    d_lo, d_hi, d_fl = d(lo), d(hi), d(flags)
    d_st.zero_()
    g.qtrim_device(d_b, None, d(offsets.astype(np.int32)), len(lo), True, d_lo, d_hi, d_fl, dev_cfg(g, **case), d_st)
    torch.cuda.synchronize()
    assert np.array_equal(d_lo.cpu().numpy(), want[0]) and np.array_equal(d_hi.cpu().numpy(), want[1])
    assert np.array_equal(d_fl.cpu().numpy(), want[2]) and d_st.cpu().tolist() == list(want[3])
