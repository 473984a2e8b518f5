"""GPU parity of Seal's multi-value table and per-pair assignment (include/seal_b200.h) against the oracle, through
the C ABI. Bit-exact: the (key, id) table, storedKmers, every per-unit output, the counters and the per-reference
totals. Reference: jgi/Seal.java:1760-1946, :2186-2276, :2386-2606, :2654-2708, :2864-2907."""
import ctypes as C

import numpy as np
import pytest
import torch

from bbtools_b200 import _lib
from bbtools_b200 import seal as PS
from oracle import seal as S
from test_seal_oracle import OPTION_SETS, make_case, pack, rand_seq

pytestmark = pytest.mark.gpu


def build_both(cfg, refs):
    g, o = PS.SealIndexGPU(cfg), S.SealOracle(cfg)
    rb, ro = pack(refs)
    g.add_ref(rb, ro)
    o.add_ref(rb, ro)
    gv, ov = g.finalize(), o.finalize()
    assert gv == ov, (gv, ov)
    gk, gi = g.table()
    ok, oi = o.table()
    assert np.array_equal(gk, ok) and np.array_equal(gi, oi), "tables differ"
    return g, o


def same_results(g, o, reads, paired, first_id=0):
    b, off = pack(reads)
    gr, gs = g.process(b, off, paired, first_id)
    wr, ws = o.process(b, off, paired, first_id)
    for name, x in wr.fields().items():
        y = gr.fields()[name]
        if not np.array_equal(x, y):
            bad = np.flatnonzero(x != y)[:5]
            raise AssertionError(f"{name} differs at {bad}: want {x[bad]}, got {y[bad]}")
    assert gs.as_dict() == ws.as_dict()
    for a, c in zip(g.scaffold_counts(), o.scaffold_counts()):
        assert np.array_equal(a, c), "per-reference totals differ"
    return gr, gs


@pytest.mark.parametrize("i", range(len(OPTION_SETS)))
def test_option_sets(i):
    cfg = PS.make_cfg(**OPTION_SETS[i])
    small = cfg.hdist == 2
    refs, reads = make_case(300 + i, n_refs=4 if small else 8, ref_len=150 if small else 400, n_frag=400, k=cfg.k, read_len=150)
    g, o = build_both(cfg, refs)
    _, st = same_results(g, o, reads, True, first_id=11)
    assert st.reads_matched > 0
    _, reads1 = make_case(400 + i, n_refs=4 if small else 8, ref_len=150 if small else 400, n_frag=300, paired=False, k=cfg.k)
    same_results(g, o, reads1, False, first_id=5)  # totals accumulate over calls on both sides
    assert g.launches > 0


@pytest.mark.parametrize("kw", [dict(), dict(clearzone_fraction=0.02, ambig_mode=PS.AMBIG_ALL), dict(restrict_right=300, k=27),
                                dict(restrict_left=280, match_mode=PS.MATCH_UNIQUE), dict(keep_pairs_together=0, hdist=1, k=25)])
def test_long_reads_cross_chunks(kw):
    cfg = PS.make_cfg(**kw)
    refs, reads = make_case(77, n_refs=6, ref_len=1500, n_frag=120, k=cfg.k, read_len=900, n_rate=0.004)
    g, o = build_both(cfg, refs)
    same_results(g, o, reads, True)


def test_many_ids_per_pair_spill_to_global_lists():
    rng = np.random.default_rng(3)
    shared = rand_seq(rng, 60)
    refs = [rand_seq(rng, 40) + shared + rand_seq(rng, 40) for _ in range(300)]
    reads = []
    for j in range(40):
        if j % 2 == 0:
            src = refs[int(rng.integers(0, 300))]
            reads += [src[20:130], shared + rand_seq(rng, 30)]
        else:  # every reference ties
            reads += [shared, shared[5:]]
    for kw in (dict(ambig_mode=PS.AMBIG_ALL, ids_stride=320), dict(), dict(ambig_mode=PS.AMBIG_FIRST),
               dict(ambig_mode=PS.AMBIG_ALL, keep_pairs_together=0, clearzone=100, ids_stride=16)):
        cfg = PS.make_cfg(**kw)
        g, o = build_both(cfg, refs)
        gr, _ = same_results(g, o, reads, True, first_id=123456789012)
        assert gr.n_sites.max() == 300


def test_list_overflow_is_an_error():
    rng = np.random.default_rng(4)
    shared = rand_seq(rng, 40)
    refs = [shared + rand_seq(rng, 20) for _ in range(1200)]
    g = PS.SealIndexGPU(PS.make_cfg())
    g.add_ref(*pack(refs))
    g.finalize()
    with pytest.raises(RuntimeError, match="distinct reference ids"):
        g.process(*pack([shared]), False)
    res, _ = g.process(*pack([refs[5][20:]]), False)  # the handle stays usable
    assert res.n_assigned[0] == 1 and res.first_id[0] == 6


def test_empty_inputs():
    cfg = PS.make_cfg()
    g, o = build_both(cfg, ["ACGT"])  # shorter than k: empty table
    same_results(g, o, ["ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT", ""], True)
    g, o = build_both(cfg, [rand_seq(np.random.default_rng(1), 100)])
    gr, gs = g.process(np.zeros(0, np.uint8), np.zeros(1, np.int64), False)
    assert len(gr.n_assigned) == 0 and gs.reads_in == 0
    same_results(g, o, ["", "", "ACG", "N" * 100], True)


def test_medium_batch_default_flags():
    rng = np.random.default_rng(12)
    base = rand_seq(rng, 3000)
    refs = []
    for r in range(400):
        if r % 4 == 0:  # strains: copies of one genome piece with a few substitutions
            p = int(rng.integers(0, 1000))
            s = list(base[p:p + 2000])
            for q in rng.integers(0, 2000, 20):
                s[int(q)] = "ACGT"[int(rng.integers(0, 4))]
            refs.append("".join(s))
        else:
            refs.append(rand_seq(rng, 2000))
    ra = [np.frombuffer(r.encode(), np.uint8) for r in refs]
    comp = np.zeros(256, np.uint8)
    for a, c in zip(b"ACGTN", b"TGCAN"):
        comp[a] = c
    n = 40000
    src = rng.integers(0, 400, n)
    pos = rng.integers(0, 2000 - 150, n)
    mat = np.stack([ra[s][p:p + 150] for s, p in zip(src, pos)])
    err = rng.random((n, 150)) < 0.01
    mat = np.where(err, np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, (n, 150))], mat)
    flip = rng.integers(0, 2, n).astype(bool)
    mat[flip] = comp[mat[flip]][:, ::-1]
    reads = [bytes(row).decode() for row in mat]
    for kw in (dict(), dict(ambig_mode=PS.AMBIG_ALL, clearzone=10), dict(keep_pairs_together=0, ambig_mode=PS.AMBIG_TOSS)):
        cfg = PS.make_cfg(**kw)
        g, o = build_both(cfg, refs)
        gr, gs = same_results(g, o, reads, True, first_id=1000)
        assert gs.reads_matched > 30000


def test_device_entry_point():
    cfg = PS.make_cfg(ambig_mode=PS.AMBIG_ALL)
    refs, reads = make_case(9, n_refs=8, ref_len=500, n_frag=2000, k=31, read_len=150)
    g, o = build_both(cfg, refs)
    b, off = pack(reads)
    want, wst = o.process(b, off, True, 42)
    lib = _lib.load()
    dev = torch.device("cuda:0")
    d_b = torch.from_numpy(np.concatenate([b, np.zeros(16, np.uint8)])).to(dev)
    d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
    nu = len(reads) // 2
    stride = cfg.ids_stride
    d_res = torch.zeros(nu * (4 + stride), dtype=torch.int32, device=dev)
    d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
    out = PS.SealOut()
    base = d_res.data_ptr()
    out.n_assigned, out.first_id, out.n_sites, out.max_hits = base, base + 4 * nu, base + 8 * nu, base + 12 * nu
    out.ids = base + 16 * nu
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):  # the counters are added to
        rc = lib.seal_b200_process_device(g.h, d_b.data_ptr(), d_off.data_ptr(), len(reads), 1, 42, C.byref(out), d_stats.data_ptr(), st)
        assert rc == 0, lib.seal_b200_last_error(g.h)
    torch.cuda.synchronize()
    r = d_res.cpu().numpy()
    assert np.array_equal(r[:nu], want.n_assigned) and np.array_equal(r[nu:2 * nu], want.first_id)
    assert np.array_equal(r[2 * nu:3 * nu], want.n_sites) and np.array_equal(r[3 * nu:4 * nu], want.max_hits)
    assert np.array_equal(r[4 * nu:], want.ids)
    s = d_stats.cpu().numpy()
    assert list(s[:6] // 2) == list(wst.as_dict().values()) and not (s[:6] % 2).any()


def test_rejects_unsupported_flags():
    lib = _lib.load()
    for kw, msg in ((dict(k=32), "K must range"), (dict(hdist=3), "hdist"), (dict(min_kmer_hits=0), "minKmerHits"),
                    (dict(k=3, mid_mask_len=2), "Middle-masking")):
        with pytest.raises(ValueError, match=msg):
            PS.SealIndexGPU(PS.make_cfg(**kw))
    g = PS.SealIndexGPU(PS.make_cfg())
    with pytest.raises(RuntimeError, match="before finalize"):
        g.process(*pack(["ACGT"]), False)


def test_front_end_files_equal_the_oracle_engine(tmp_path):
    """seal.sh from files on the GPU engine against the same front end on the oracle engine: byte-identical outputs"""
    from test_seal_host_cpu import _write_inputs
    refs, reads = make_case(41, n_refs=8, ref_len=400, n_frag=600, k=31, read_len=140)
    reads = [r if len(r) > 0 else "A" for r in reads]
    ref, r1, r2 = _write_inputs(tmp_path, refs, reads)
    outs = {}
    for tag, engine in (("gpu", None), ("ora", S.SealOracle)):
        args = [f"in={r1}", f"in2={r2}", f"ref={ref}", f"outm={tmp_path}/{tag}_m.fq", f"outu={tmp_path}/{tag}_u1.fq",
                f"outu2={tmp_path}/{tag}_u2.fq", f"stats={tmp_path}/{tag}_stats.txt", "ambig=random", "cz=2"]
        tool = PS.Seal(args, engine=engine)
        outs[tag] = (tool.process(block_bytes=20000), (tool.stored, tool.entries, tool.ref_kmers))
    assert outs["gpu"] == outs["ora"] and outs["gpu"][0]["reads_matched"] > 0
    for name in ("m.fq", "u1.fq", "u2.fq", "stats.txt"):
        assert open(f"{tmp_path}/gpu_{name}", "rb").read() == open(f"{tmp_path}/ora_{name}", "rb").read(), name
