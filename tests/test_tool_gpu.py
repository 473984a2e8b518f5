"""The reads-in / reads-out surface (bbtools_b200.bbduk.BBDuk, the bbduk.sh command line) on FASTQ files: native
feed vs the plain-Python feed, and both against records cut with the ORACLE's coordinates."""
import os

import numpy as np
import pytest

from bbtools_b200 import F_REMOVED, make_cfg, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def write_fastq(path, bases, offsets, first, step, tag):
    with open(path, "wb") as f:
        for i in range(first, len(offsets) - 1, step):
            s = bytes(bases[offsets[i]:offsets[i + 1]])
            f.write(b"@pair%d %s\n" % (i // 2, tag) + s + b"\n+\n" + b"I" * len(s) + b"\n")


@pytest.mark.parametrize("extra", [[], ["ottm=t"], ["rieb=f", "minlen=60"]])
def test_bbduk_tool_paired_fastq(tmp_path, extra):
    from bbtools_b200.bbduk import BBDuk
    from bbtools_b200.fasta import read_fasta
    from oracle.oracle import Oracle
    bases, offsets = synth.paired_adapter_reads(4000, seed=41)
    r1, r2 = tmp_path / "r1.fq", tmp_path / "r2.fq"
    write_fastq(r1, bases, offsets, 0, 2, b"1:N:0")
    write_fastq(r2, bases, offsets, 1, 2, b"2:N:0")
    common = [f"in={r1}", f"in2={r2}", f"ref={GOLDEN}/adapters.fa", "ktrim=r", "k=23", "mink=11", "hdist=1", "tpe"] + extra
    outs = {}
    for mode in ("native", "python"):
        o1, o2, m1, m2 = (tmp_path / f"{mode}_{x}.fq" for x in ("o1", "o2", "m1", "m2"))
        tool = BBDuk(common + [f"out={o1}", f"out2={o2}", f"outm={m1}", f"outm2={m2}", f"stats={tmp_path}/{mode}.stats"])
        st = tool.process(native=(mode == "native"))
        outs[mode] = tuple(open(p, "rb").read() for p in (o1, o2, m1, m2)) + (open(f"{tmp_path}/{mode}.stats").read(),)
        cfg = tool.cfg
    assert outs["native"] == outs["python"]
    # the same files from the oracle's coordinates
    _, rb, roff = read_fasta(os.path.join(GOLDEN, "adapters.fa"))
    ora = Oracle(cfg)
    ora.add_ref(rb, roff)
    ora.finalize()
    want, wst = ora.process(bases, offsets, True)
    assert wst.as_dict() == st.as_dict()
    ottm = "ottm=t" in extra
    exp = [[], [], [], []]
    for i in range(len(offsets) - 1):
        rem = bool(want.flags[i & ~1] & F_REMOVED)
        s = bytes(bases[offsets[i]:offsets[i + 1]])
        a, b = (int(want.lo[i]), int(want.hi[i])) if (not rem or ottm) else (0, len(s))
        recd = b"@pair%d %s\n" % (i // 2, b"1:N:0" if i % 2 == 0 else b"2:N:0") + s[a:b] + b"\n+\n" + b"I" * (b - a) + b"\n"
        exp[(2 if rem else 0) + (i & 1)].append(recd)
    for k in range(4):
        assert outs["native"][k] == b"".join(exp[k]), k
    assert len(outs["native"][2]) > 0  # some pairs were removed


@pytest.mark.parametrize("strict", [True, False])
def test_bbduk_tool_canonical_adapter_trimming_with_tbo(tmp_path, strict):
    """bbduk.sh in=.. in2=.. ref=adapters ktrim=r k=23 mink=11 hdist=1 tpe tbo: k-mer block, then trim by overlap"""
    from bbtools_b200.bbduk import BBDuk
    from bbtools_b200.fasta import read_fasta
    from oracle import tbo as otbo
    from oracle.oracle import Oracle
    bases, offsets = synth.paired_adapter_reads(6000, seed=43)
    rng = np.random.default_rng(2)
    quals = (33 + rng.integers(12, 41, len(bases))).astype(np.uint8)
    noisy = rng.random(len(offsets) - 1) < 0.2  # reads whose expected errors trip strictoverlap's filter
    for i in np.nonzero(noisy)[0]:
        quals[offsets[i]:offsets[i + 1]] = 33 + rng.integers(2, 12, offsets[i + 1] - offsets[i])
    paths = []
    for first, tag in ((0, b"1:N:0"), (1, b"2:N:0")):
        path = tmp_path / f"r{first + 1}.fq"
        with open(path, "wb") as f:
            for i in range(first, len(offsets) - 1, 2):
                f.write(b"@pair%d %s\n" % (i // 2, tag) + bytes(bases[offsets[i]:offsets[i + 1]]) + b"\n+\n" +
                        bytes(quals[offsets[i]:offsets[i + 1]]) + b"\n")
        paths.append(path)
    common = [f"in={paths[0]}", f"in2={paths[1]}", f"ref={GOLDEN}/adapters.fa", "ktrim=r", "k=23", "mink=11", "hdist=1", "tpe",
              "tbo", f"strictoverlap={'t' if strict else 'f'}"]
    outs = {}
    for mode in ("native", "python"):
        o1, o2, m1 = (tmp_path / f"{mode}_{x}.fq" for x in ("o1", "o2", "m1"))
        tool = BBDuk(common + [f"out={o1}", f"out2={o2}", f"outm={m1}"])
        tool.process(native=(mode == "native"))
        outs[mode] = tuple(open(p, "rb").read() for p in (o1, o2, m1)) + (list(tool.tbo_stats),)
        cfg = tool.cfg
    assert outs["native"] == outs["python"]
    _, rb, roff = read_fasta(os.path.join(GOLDEN, "adapters.fa"))
    ora = Oracle(cfg)
    ora.add_ref(rb, roff)
    ora.finalize()
    want, _ = ora.process(bases, offsets, True)
    whi, _, _, wst = otbo.process(bases, quals, offsets, want.lo, want.hi, want.flags, otbo.default_params(strict))
    assert list(wst) == outs["native"][3]
    assert wst[0] > 100
    exp = [[], []]
    for i in range(len(offsets) - 1):
        if want.flags[i & ~1] & F_REMOVED:
            continue
        a, b = int(want.lo[i]), int(whi[i])
        exp[i & 1].append(b"@pair%d %s\n" % (i // 2, b"1:N:0" if i % 2 == 0 else b"2:N:0") +
                          bytes(bases[offsets[i] + a:offsets[i] + b]) + b"\n+\n" + bytes(quals[offsets[i] + a:offsets[i] + b]) + b"\n")
    assert outs["native"][0] == b"".join(exp[0]) and outs["native"][1] == b"".join(exp[1])
    if strict:
        assert np.any(whi != want.hi)


def test_bbduk_tool_whole_chain(tmp_path):
    """ktrim=r ... tpe tbo trimpolya=4 qtrim=rl trimq=10 maxns=1 entropy=0.6: every device step of the per-pair loop in the
    reference's order (k-mer block, tbo, poly-X, quality trimming, quality / N filters, entropy filter)"""
    from bbtools_b200.bbduk import BBDuk
    from bbtools_b200.fasta import read_fasta
    from oracle import entropy as oe
    from oracle import qtrim as oq
    from oracle import tbo as otbo
    from oracle.oracle import Oracle
    bases, offsets = synth.paired_adapter_reads(5000, seed=47)
    bases = bases.copy()
    rng = np.random.default_rng(6)
    n = len(offsets) - 1
    for i in np.nonzero(rng.random(n) < 0.15)[0]:  # low-complexity reads, poly-A tails
        if rng.random() < 0.5:
            unit = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rng.integers(1, 4)))]
            bases[offsets[i]:offsets[i + 1]] = np.tile(unit, 150)[:150]
        else:
            bases[offsets[i + 1] - int(rng.integers(4, 30)):offsets[i + 1]] = ord("A")
    pos = np.arange(len(bases)) % 150
    quals = (np.clip(40 - (pos * rng.integers(0, 45, len(bases))) // 150 + rng.integers(-3, 4, len(bases)), 2, 41) + 33).astype(np.uint8)
    paths = []
    for first, tag in ((0, b"1:N:0"), (1, b"2:N:0")):
        path = tmp_path / f"r{first + 1}.fq"
        with open(path, "wb") as f:
            for i in range(first, n, 2):
                f.write(b"@pair%d %s\n" % (i // 2, tag) + bytes(bases[offsets[i]:offsets[i + 1]]) + b"\n+\n" +
                        bytes(quals[offsets[i]:offsets[i + 1]]) + b"\n")
        paths.append(path)
    common = [f"in={paths[0]}", f"in2={paths[1]}", f"ref={GOLDEN}/adapters.fa", "ktrim=r", "k=23", "mink=11", "hdist=1", "tpe", "tbo",
              "trimpolya=4", "qtrim=rl", "trimq=10", "maxns=1", "entropy=0.6", "minlen=25"]
    outs = {}
    for mode in ("native", "python"):
        o1, o2, m1 = (tmp_path / f"{mode}_{x}.fq" for x in ("o1", "o2", "m1"))
        tool = BBDuk(common + [f"out={o1}", f"out2={o2}", f"outm={m1}"])
        tool.process(native=(mode == "native"))
        outs[mode] = tuple(open(p, "rb").read() for p in (o1, o2, m1)) + (list(tool.tbo_stats), list(tool.qtrim_stats),
                                                                             list(tool.entropy_stats))
        cfg = tool.cfg
    assert outs["native"] == outs["python"]
    _, rb, roff = read_fasta(os.path.join(GOLDEN, "adapters.fa"))
    ora = Oracle(cfg)
    ora.add_ref(rb, roff)
    ora.finalize()
    want, _ = ora.process(bases, offsets, True)
    hi1, _, _, tst = otbo.process(bases, quals, offsets, want.lo, want.hi, want.flags)
    fl1 = want.flags | np.where(hi1 != want.hi, np.uint8(0x20), np.uint8(0))
    lo2, hi2, fl2, qst = oq.process(bases, quals, offsets, True, want.lo, hi1, fl1, oq.params(qtrim="rl", trimq=10.0, maxns=1, polya=4, minlen=25))
    hi3, fl3, est = oe.process(bases, offsets, True, lo2, hi2, fl2, oe.params(cutoff=0.6))
    assert outs["native"][3:] == (list(tst), list(qst), list(est))
    assert tst[0] > 50 and qst[0] > 1000 and qst[6] > 100 and est[0] > 100
    exp = [[], []]
    for i in range(n):
        if fl3[i & ~1] & F_REMOVED:
            continue
        a, b = int(lo2[i]), int(hi3[i])
        exp[i & 1].append(b"@pair%d %s\n" % (i // 2, b"1:N:0" if i % 2 == 0 else b"2:N:0") +
                          bytes(bases[offsets[i] + a:offsets[i] + b]) + b"\n+\n" + bytes(quals[offsets[i] + a:offsets[i] + b]) + b"\n")
    assert outs["native"][0] == b"".join(exp[0]) and outs["native"][1] == b"".join(exp[1])


def test_bbduk_tool_single_kfilter(tmp_path):
    from bbtools_b200.bbduk import BBDuk
    ref = synth.random_reference(2, 50_000, seed=7)
    rf = tmp_path / "ref.fa"
    with open(rf, "wb") as f:
        for i in range(2):
            f.write(b">scaf%d\n" % i + bytes(ref[0][ref[1][i]:ref[1][i + 1]]) + b"\n")
    bases, offsets = synth.contaminant_reads(5000, ref[0], seed=3, contam_pct=25)
    rq = tmp_path / "r.fq"
    write_fastq(rq, bases, offsets, 0, 1, b"se")
    res = {}
    for mode in ("native", "python"):
        o, m = tmp_path / f"{mode}_clean.fq", tmp_path / f"{mode}_contam.fq"
        st = BBDuk([f"in={rq}", f"ref={rf}", "k=31", f"out={o}", f"outm={m}"]).process(native=(mode == "native"))
        res[mode] = (open(o, "rb").read(), open(m, "rb").read(), st.as_dict())
    assert res["native"] == res["python"]
    assert 1000 < res["native"][2]["reads_kfiltered"] < 1500
    assert res["native"][0].count(b"\n") + res["native"][1].count(b"\n") == 4 * 5000


@pytest.mark.parametrize("flag,mode", [("entropymask", 1), ("entropymask=lc", 2), ("entropytrim=rl", 3)])
def test_bbduk_tool_entropy_mask_and_trim(tmp_path, flag, mode):
    """ktrim=r ... entropy=0.6 entropymask / entropytrim: the k-mer block, then low-entropy windows masked (to N with quality 0, or
    to lower case) or trimmed from both ends (jgi/BBDuk.java:3055-3067); records and counters against the oracles"""
    from bbtools_b200.bbduk import BBDuk
    from bbtools_b200.fasta import read_fasta
    from oracle import entropy as oe
    from oracle.oracle import Oracle
    from test_entropy_oracle import entropy_batch
    bases, offsets, _, _, _ = entropy_batch(3000, 77, L=140)
    keep = np.diff(offsets) > 0
    seqs = [bytes(bases[offsets[i]:offsets[i + 1]]) for i in range(len(keep)) if keep[i]]
    bases = np.frombuffer(b"".join(seqs), np.uint8).copy()
    offsets = np.zeros(len(seqs) + 1, np.int64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    path = tmp_path / "r.fq"
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    o1, m1 = tmp_path / "o.fq", tmp_path / "m.fq"
    tool = BBDuk([f"in={path}", f"ref={GOLDEN}/adapters.fa", "ktrim=r", "k=23", "mink=11", "hdist=1", "entropy=0.6", flag, "minlen=1",
                  f"out={o1}", f"outm={m1}"])
    tool.process()
    _, rb, roff = read_fasta(os.path.join(GOLDEN, "adapters.fa"))
    ora = Oracle(tool.cfg)
    ora.add_ref(rb, roff)
    ora.finalize()
    want, _ = ora.process(bases, offsets, False)
    wlo, whi, bits, moff, wst = oe.mask(bases, offsets, False, want.lo, want.hi, want.flags, oe.params(cutoff=0.6, rieb=True), mode)
    assert list(tool.entropy_stats) == list(wst) and wst[1] > 200
    exp = [[], []]
    for i, s in enumerate(seqs):
        s, q = bytearray(s), bytearray(b"I" * len(s))
        for j in range(int(want.hi[i]) - int(want.lo[i])):
            if (int(bits[moff[i] + (j >> 5)]) >> (j & 31)) & 1:
                p = int(want.lo[i]) + j
                if mode == 1 and s[p] != ord("N"):
                    s[p], q[p] = ord("N"), 33
                elif mode == 2:
                    s[p:p + 1] = bytes(s[p:p + 1]).lower()
        rem = bool(want.flags[i] & F_REMOVED)
        a, b = (0, len(s)) if rem else (int(wlo[i]), int(whi[i]))
        exp[1 if rem else 0].append(b"@r%d\n" % i + bytes(s[a:b]) + b"\n+\n" + bytes(q[a:b]) + b"\n")
    assert open(o1, "rb").read() == b"".join(exp[0])
    assert (open(m1, "rb").read() if os.path.exists(m1) else b"") == b"".join(exp[1])


def test_bbduk_tool_streams_the_input_in_bounded_blocks(tmp_path, monkeypatch):
    """the native feed with blocks far smaller than the files (one gzipped): the same outputs, counters and stats file as
    the plain-Python feed that holds everything at once; the k-mer block followed by tbo + quality trimming"""
    import gzip

    from bbtools_b200.bbduk import BBDuk
    bases, offsets = synth.paired_adapter_reads(6000, seed=43)
    r1, r2 = tmp_path / "r1.fq.gz", tmp_path / "r2.fq"
    write_fastq(tmp_path / "r1.fq", bases, offsets, 0, 2, b"1:N:0 a longer header so that the two files drift apart")
    write_fastq(r2, bases, offsets, 1, 2, b"2")
    with gzip.open(r1, "wb") as f:
        f.write(open(tmp_path / "r1.fq", "rb").read())
    for extra in ([], ["tbo", "qtrim=rl", "trimq=10", "minlen=30"]):
        common = [f"in={r1}", f"in2={r2}", f"ref={GOLDEN}/adapters.fa", "ktrim=r", "k=23", "mink=11", "hdist=1", "tpe"] + extra
        outs = {}
        for mode in ("native", "python"):
            o1, o2, m1, m2 = (tmp_path / f"{mode}_{x}.fq" for x in ("o1", "o2", "m1", "m2"))
            tool = BBDuk(common + [f"out={o1}", f"out2={o2}", f"outm={m1}", f"outm2={m2}", f"stats={tmp_path}/{mode}.stats"])
            if mode == "native":
                monkeypatch.setenv("BBDUK_B200_FEED_BLOCK", "50000")  # ~40 blocks
            else:
                monkeypatch.delenv("BBDUK_B200_FEED_BLOCK", raising=False)
            st = tool.process(native=(mode == "native"))
            outs[mode] = tuple(open(p, "rb").read() for p in (o1, o2, m1, m2)) + (open(f"{tmp_path}/{mode}.stats").read(), st.as_dict(),
                                                                                   None if tool.tbo_stats is None else list(tool.tbo_stats),
                                                                                   None if tool.qtrim_stats is None else list(tool.qtrim_stats))
        assert outs["native"] == outs["python"], extra
        assert len(outs["native"][0]) > 100000 and len(outs["native"][2]) > 0
