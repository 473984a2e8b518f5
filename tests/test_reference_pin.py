"""Holds the oracle (CPU) and the CUDA path (GPU) to digests of what the REAL reference wrote for the same command lines
and seeded inputs (tools/pin_reference.py). While tests/golden/reference_digests.json does not exist -- no JRE in this
image or on the GPU boxes, profiles/r02_java_probe.txt -- the tests xfail with "parity unpinned": they are the consumer
that is waiting for the file, not a skip."""
import json
import os

import numpy as np
import pytest

import pin_common as pc


def _digests():
    if not os.path.exists(pc.DIGESTS):
        pytest.xfail("PARITY UNPINNED: no digests of the real reference (run tools/pin_reference.sh where a JRE exists)")
    return json.load(open(pc.DIGESTS))


def _cfg(flags, generation):
    from bbtools_b200.bbduk import parse_args
    cfg, _ = parse_args([f"ref={pc.GOLDEN}/adapters.fa"] + flags, generation)
    return cfg


def _check(case, cls, want, rendered, stored):
    keys = ["out", "out2", "outm", "outm2"]
    assert want.get("added_kmers") in (None, stored), (case, cls, "Added N kmers")
    for key, data in zip(keys, rendered):
        if want.get(key) is not None:
            assert pc.sha(data) == want[key], (case, cls, key)


def test_pin_inputs_are_seeded():
    """the recipe and the tests regenerate identical inputs (no files travel)"""
    b1, o1, _ = pc.inputs("pairs")
    b2, o2, _ = pc.inputs("pairs")
    assert np.array_equal(b1, b2) and np.array_equal(o1, o2)
    assert pc.sha(bytes(b1[:150])) == pc.sha(bytes(b2[:150]))
    for name, (kind, flags) in pc.CASES.items():
        for gen in (0, 1):
            _cfg(flags, gen)  # every pinned command line parses


@pytest.mark.parametrize("case", sorted(pc.CASES))
def test_oracle_matches_reference_digests(case):
    from bbtools_b200.fasta import read_fasta
    from oracle.oracle import Oracle
    doc = _digests()
    kind, flags = pc.CASES[case]
    bases, offsets, paired = pc.inputs(kind)
    _, rb, roff = read_fasta(os.path.join(pc.GOLDEN, "adapters.fa"))
    for gen, cls in enumerate(pc.MAIN_CLASSES):
        want = doc["cases"][case][cls]
        if "error" in want:
            continue
        o = Oracle(_cfg(flags, gen))
        o.add_ref(rb, roff)
        stored = o.finalize()
        mask = "ktrim=N" in flags
        out, _ = o.process(bases, offsets, paired, want_mask=mask)
        rendered = pc.render(bases, offsets, paired, out.lo, out.hi, out.flags, out.maskbits if mask else None,
                             out.mask_off if mask else None)
        _check(case, cls, want, rendered, stored)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(pc.CASES))
def test_gpu_matches_reference_digests(case):
    from bbtools_b200.bbduk import BBDukIndexGPU
    from bbtools_b200.fasta import read_fasta
    doc = _digests()
    kind, flags = pc.CASES[case]
    bases, offsets, paired = pc.inputs(kind)
    _, rb, roff = read_fasta(os.path.join(pc.GOLDEN, "adapters.fa"))
    for gen, cls in enumerate(pc.MAIN_CLASSES):
        want = doc["cases"][case][cls]
        if "error" in want:
            continue
        g = BBDukIndexGPU(_cfg(flags, gen))
        g.add_ref(rb, roff)
        stored = g.finalize()
        mask = "ktrim=N" in flags
        out, _ = g.process(bases, offsets, paired, want_mask=mask)
        rendered = pc.render(bases, offsets, paired, out.lo, out.hi, out.flags, out.maskbits if mask else None,
                             out.mask_off if mask else None)
        _check(case, cls, want, rendered, stored)


# ---- Seal ----------------------------------------------------------------------------------------------------------------
def _seal_engine_outputs(engine_cls, flags):
    """(rendered outm/outm2/outu/outu2, stats lines without the file header, storedKmers) of one pinned Seal command line"""
    from bbtools_b200 import seal as PS
    cfg, _ = PS.parse_seal_args(flags)
    eng = engine_cls(cfg)
    names = []
    for name, s in pc.seal_refs():
        names.append(name)
        b = np.frombuffer(s, np.uint8)
        eng.add_ref(b, np.array([0, len(b)], np.int64))
    stored = eng.finalize()[0]
    bases, offsets = pc.seal_reads()
    res, st = eng.process(bases, offsets, True, 0)
    hit = (res.n_assigned[0::2] + res.n_assigned[1::2]) > 0 if not cfg.keep_pairs_together else res.n_assigned > 0
    stats = PS.format_stats(names, st.as_dict(), eng.scaffold_counts(), "x")
    stats = "".join(ln + "\n" for ln in stats.splitlines() if not ln.startswith("#File"))
    return pc.render_seal(bases, offsets, hit), stats, stored


def _check_seal(case, want, got):
    rendered, stats, stored = got
    assert want.get("added_kmers") in (None, stored), (case, "Added N kmers")
    for key, data in zip(("outm", "outm2", "outu", "outu2"), rendered):
        if want.get(key) is not None:
            assert pc.sha(data) == want[key], (case, key)
    if want.get("stats") is not None:
        assert pc.sha(stats.encode()) == want["stats"], (case, "stats")


def _seal_digests():
    doc = _digests()
    if "seal_cases" not in doc:
        pytest.xfail("PARITY UNPINNED: reference_digests.json holds no Seal cases (rerun tools/pin_reference.sh)")
    return doc["seal_cases"]


def test_seal_pin_inputs_are_seeded_and_parse():
    from bbtools_b200 import seal as PS
    b1, o1 = pc.seal_reads(200)
    b2, o2 = pc.seal_reads(200)
    assert np.array_equal(b1, b2) and np.array_equal(o1, o2) and pc.seal_refs() == pc.seal_refs()
    for flags in pc.SEAL_CASES.values():
        PS.parse_seal_args(flags)


def test_seal_pin_consumer_works_on_digests_of_the_oracle_itself(tmp_path, monkeypatch):
    """NOT a pin: the consumer below is run against digests made from the oracle's own output, so that it is known to work
    the day real digests arrive (a wrong byte in them must fail it: checked too)."""
    from oracle import seal as S
    case = "seal_ambig_all_cz3"
    rendered, stats, stored = _seal_engine_outputs(S.SealOracle, pc.SEAL_CASES[case])
    assert all(len(x) > 0 for x in rendered)
    want = {"added_kmers": stored, "stats": pc.sha(stats.encode())}
    want.update({k: pc.sha(d) for k, d in zip(("outm", "outm2", "outu", "outu2"), rendered)})
    _check_seal(case, want, (rendered, stats, stored))
    want["outu"] = pc.sha(rendered[2] + b"x")
    with pytest.raises(AssertionError):
        _check_seal(case, want, (rendered, stats, stored))


@pytest.mark.parametrize("case", sorted(pc.SEAL_CASES))
def test_seal_oracle_matches_reference_digests(case):
    from oracle import seal as S
    want = _seal_digests()[case]
    if "error" in want:
        pytest.skip("the reference failed on this command line: " + want["error"][-200:])
    _check_seal(case, want, _seal_engine_outputs(S.SealOracle, pc.SEAL_CASES[case]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(pc.SEAL_CASES))
def test_seal_gpu_matches_reference_digests(case):
    from bbtools_b200 import seal as PS
    want = _seal_digests()[case]
    if "error" in want:
        pytest.skip("the reference failed on this command line: " + want["error"][-200:])
    _check_seal(case, want, _seal_engine_outputs(PS.SealIndexGPU, pc.SEAL_CASES[case]))
