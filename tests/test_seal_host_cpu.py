"""CPU tests of the Seal host side: seal.sh's flag surface (jgi/Seal.java:150-380) and the N>1 logic (world_size-2 gloo,
the oracle standing in for the per-GPU engine): sharded == unsharded, including ambig=random, which depends on the
pair's global numericID."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from bbtools_b200 import seal as PS
from test_seal_oracle import make_case, pack


def test_defaults_match_seal():
    c, files = PS.parse_seal_args(["in=r.fq", "ref=a.fa,b.fa", "out=m.fq"])
    d = PS.make_cfg()
    for name, _ in PS.SealCfg._fields_:
        if name != "reserved":
            assert getattr(c, name) == getattr(d, name), name
    assert (c.k, c.rcomp, c.mask_middle, c.hdist, c.ambig_mode, c.match_mode, c.keep_pairs_together, c.min_kmer_hits) == \
        (31, 1, 1, 0, PS.AMBIG_RANDOM, PS.MATCH_ALL, 1, 1)
    assert files == {"in1": "r.fq", "ref": "a.fa,b.fa", "outm1": "m.fq"}  # out= is outm= (jgi/Seal.java:177)


@pytest.mark.parametrize("args,want", [
    (["k=25", "hammingdistance=1", "mm=f"], dict(k=25, hdist=1, mask_middle=0)),
    (["mm=3"], dict(mask_middle=1, mid_mask_len=3)),
    (["mm=0"], dict(mask_middle=0, mid_mask_len=0)),
    (["maskmiddle"], dict(mask_middle=1)),
    (["ambig=toss", "match=unique"], dict(ambig_mode=PS.AMBIG_TOSS, match_mode=PS.MATCH_UNIQUE)),
    (["ambiguous=Best", "mode=first"], dict(ambig_mode=PS.AMBIG_FIRST, match_mode=PS.MATCH_FIRST)),
    (["ambig=all", "fbm=f"], dict(ambig_mode=PS.AMBIG_ALL, match_mode=PS.MATCH_FIRST)),
    (["fum"], dict(match_mode=PS.MATCH_UNIQUE)),
    (["kpt=f", "cz=7", "mkh=3", "mkf=0.25"], dict(keep_pairs_together=0, clearzone=7, min_kmer_hits=3, min_kmer_fraction=0.25)),
    (["clearzone=0.1"], dict(clearzone=0, clearzone_fraction=np.float32(0.1))),
    (["czf=0.05", "rskip=3", "qskip=2", "speed=4"], dict(clearzone_fraction=np.float32(0.05), rskip=3, qskip=2, speed=4)),
    (["restrictleft=50", "restrictright=40", "fn", "rcomp=f"], dict(restrict_left=50, restrict_right=40, forbid_ns=1, rcomp=0)),
    (["edist=0", "qhdist=0", "arrayhf=t", "ordered", "prealloc=0.5"], dict()),
])
def test_flags(args, want):
    c, _ = PS.parse_seal_args(args)
    d = PS.make_cfg(**{k: (float(v) if isinstance(v, np.floating) else v) for k, v in want.items()})
    for name, _ in PS.SealCfg._fields_:
        if name != "reserved":
            assert getattr(c, name) == getattr(d, name), name


@pytest.mark.parametrize("args", [["edist=1"], ["qhdist=2"], ["processcontainedref=t"], ["countvector"], ["ambig=sometimes"],
                                  ["match"], ["k=32"], ["hdist=4"], ["speed=17"], ["nonsense=1"]])
def test_rejected_flags(args):
    with pytest.raises(ValueError):
        PS.parse_seal_args(args)


def _worker(rank, world, port, kw, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import seal as S  # the checker stands in for the per-GPU engine in this CPU test
    cfg = PS.make_cfg(**kw)
    refs, reads = make_case(21, n_refs=8, ref_len=400, n_frag=301, k=cfg.k, read_len=120)
    o = S.SealOracle(cfg)
    o.add_ref(*pack(refs))
    o.finalize()
    b, off = pack(reads)
    merged, total, counts = PS.process_sharded(o, b, off, True, first_numeric_id=1000)
    if rank == 0:
        q.put((merged, total, counts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kw", [dict(), dict(keep_pairs_together=0, ambig_mode=PS.AMBIG_ALL, clearzone=4)])
def test_two_ranks_equal_one(kw):
    from oracle import seal as S
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kw, q)) for r in range(2)]
    [p.start() for p in procs]
    merged, total, counts = q.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    cfg = PS.make_cfg(**kw)
    refs, reads = make_case(21, n_refs=8, ref_len=400, n_frag=301, k=cfg.k, read_len=120)
    o = S.SealOracle(cfg)
    o.add_ref(*pack(refs))
    o.finalize()
    want, wst = o.process(*pack(reads), True, 1000)
    for name, x in want.fields().items():
        assert np.array_equal(merged[name], x), name
    assert total == wst.as_dict()
    for a, c in zip(counts, o.scaffold_counts()):
        assert np.array_equal(a, c)
    assert (want.n_sites > 1).any()  # ambiguous pairs are in the batch: ambig=random used the global pair index


def test_stats_file_format():
    names = ["zeta", "alpha", "beta", "gamma"]
    counts = [np.array([0, 4, 10, 4, 0]), np.array([0, 600, 1500, 600, 0]), np.array([0, 2, 5, 2, 0]), np.array([0, 1, 0, 2, 0])]
    st = dict(reads_in=40, bases_in=6000, reads_matched=18, bases_matched=2700, reads_unmatched=22, bases_unmatched=3300)
    txt = PS.format_stats(names, st, counts, "r1.fq", "r2.fq")
    assert txt == ("#File\tr1.fq\tr2.fq\n#Total\t40\t6000\n#Matched\t18\t45.00000%\t2700\n"
                   "#Name\tReads\tReadsPct\tBases\tBasesPct\tAmbigReads\n"
                   "alpha\t10\t25.00000%\t1500\t25.00000%\t0\n"
                   "beta\t4\t10.00000%\t600\t10.00000%\t2\n"      # ties on bases and reads: by name
                   "zeta\t4\t10.00000%\t600\t10.00000%\t1\n")
    txt3 = PS.format_stats(names, st, counts, "r.fq", columns=3, nonzero_only=False)
    assert txt3.splitlines()[:4] == ["#File\tr.fq", "#Total\t40", "#Matched\t18\t45.00000%", "#Name\tReads\tReadsPct"]
    assert txt3.splitlines()[-1] == "gamma\t0\t0.00000%" and len(txt3.splitlines()) == 8


def _write_inputs(tmp_path, refs, reads):
    ref = tmp_path / "ref.fa"
    ref.write_text("".join(f">ref{i} some description\n{s}\n" for i, s in enumerate(refs)))
    r1, r2 = tmp_path / "r1.fq", tmp_path / "r2.fq"
    with open(r1, "w") as a, open(r2, "w") as b:
        for i in range(0, len(reads), 2):
            for fh, s, m in ((a, reads[i], 1), (b, reads[i + 1], 2)):
                fh.write(f"@pair{i // 2}/{m}\n{s}\n+\n{'I' * len(s)}\n")
    return str(ref), str(r1), str(r2)


def _records(path):
    lines = open(path).read().split("\n")
    return [tuple(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4)]


@pytest.mark.parametrize("flags", [[], ["kpt=f", "ambig=all", "cz=3"], ["ambig=toss", "mkh=5"]])
def test_seal_front_end_streams_files(tmp_path, flags):
    """FASTQ in / FASTQ out in several blocks (the oracle standing in for the GPU engine): matched / unmatched partition,
    running numericID across blocks and the stats= file equal a one-shot run."""
    from oracle import seal as S
    refs, reads = make_case(31, n_refs=8, ref_len=400, n_frag=240, k=31, read_len=120)
    reads = [r if len(r) > 0 else "A" for r in reads]  # 4-line records need a base line; an empty read stays a corner of the array tests
    ref, r1, r2 = _write_inputs(tmp_path, refs, reads)
    args = [f"in={r1}", f"in2={r2}", f"ref={ref}", f"outm={tmp_path}/m1.fq", f"outm2={tmp_path}/m2.fq", f"outu={tmp_path}/u.fq",
            f"stats={tmp_path}/stats.txt"] + flags
    tool = PS.Seal(args, engine=S.SealOracle)
    assert tool.names[0] == "ref0 some description" and len(tool.names) == len(refs)
    total = tool.process(block_bytes=6000)  # a few dozen pairs per block
    cfg, _ = PS.parse_seal_args(flags)
    o = S.SealOracle(cfg)
    o.add_ref(*pack(refs))
    o.finalize()
    want, wst = o.process(*pack(reads), True, 0)
    assert total == wst.as_dict()
    hit = (want.n_assigned[0::2] + want.n_assigned[1::2]) > 0 if not cfg.keep_pairs_together else want.n_assigned > 0
    assert 0 < hit.sum() < len(hit)
    m1, m2, u = _records(f"{tmp_path}/m1.fq"), _records(f"{tmp_path}/m2.fq"), _records(f"{tmp_path}/u.fq")
    assert [r[0] for r in m1] == [f"@pair{i}/1" for i in np.flatnonzero(hit)]
    assert [r[0] for r in m2] == [f"@pair{i}/2" for i in np.flatnonzero(hit)]
    assert [r[0] for r in u] == [f"@pair{i}/{m}" for i in np.flatnonzero(~hit) for m in (1, 2)]  # one file: interleaved
    assert all(r[1] == reads[2 * i] for r, i in zip(m1, np.flatnonzero(hit)))
    assert all(len(r[1]) == len(r[3]) for r in m1 + m2 + u)
    names = [f"ref{i} some description" for i in range(len(refs))]
    assert open(f"{tmp_path}/stats.txt").read() == PS.format_stats(names, wst.as_dict(), o.scaffold_counts(), r1, r2)


def test_refstats_file_format():
    # two reference files: a.fa.gz with sequences 1-2 (1000 + 500 bases), dir/b.fasta with sequence 3 (2000 bases)
    counts = [np.array([0, 6, 2, 4]), np.array([0, 900, 300, 600]), np.array([0, 3, 1, 2]), np.array([0, 0, 1, 0])]
    st = dict(reads_in=20, bases_in=3000, reads_matched=12, bases_matched=1800, reads_unmatched=8, bases_unmatched=1200)
    txt = PS.format_refstats(["a.fa.gz", "dir/b.fasta"], [2, 1], np.array([1000, 500, 2000]), st, counts, "r.fq")
    mult = np.float32(1e9) / np.float32(12)
    want = ("#File\tr.fq\n#Reads\t20\n#Mapped\t12\n#References\t2\n#Name\tLength\tScaffolds\tBases\tCoverage\tReads\tRPKM\tFrags\tFPKM\tAmbigReads\n"
            "a\t1500\t2\t1200\t0.8000\t8\t%.4f\t4\t%.4f\t1\n" % (8 * float(mult) / 1500, 4 * float(mult) / 1500)
            + "b\t2000\t1\t600\t0.3000\t4\t%.4f\t2\t%.4f\t0\n" % (4 * float(mult) / 2000, 2 * float(mult) / 2000))
    assert txt == want
    assert "a\t1500\t2\t1200\t0.8000\t8\t444444.4" in txt  # 8 reads / 12 mapped / 1.5 kb = 444,444 RPKM
    assert PS.strip_to_core("/x/y/genome.v2.fna.gz") == "genome.v2" and PS.strip_to_core("plain") == "plain"
