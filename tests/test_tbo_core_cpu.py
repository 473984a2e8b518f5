"""CPU test of bbtools_b200/csrc/tbo_core.cuh -- the per-pair code the trim-by-overlap device lanes run (bit-plane packing
with SIMD-in-register classification, the 64-base register screen, the two insert loops) -- built for the host with
stride 1 (tests/tbo_core_host.cpp) and compared with the oracle. The GPU parity tests (test_tbo_gpu.py) check the same
code inside the kernel; this one makes its logic testable where there is no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bbtools_b200 import synth
from oracle import tbo as otbo
from test_tbo_oracle import small_pairs

HERE = os.path.dirname(os.path.abspath(__file__))
COMP, _ = otbo.tables()


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "tbo_core_host.cpp")
    hdr = os.path.join(HERE, "..", "bbtools_b200", "csrc", "tbo_core.cuh")
    out = os.path.join(HERE, "stubs", "_tbo_core_host.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", src, "-o", out])
    lib = C.CDLL(out)
    lib.tbo_host_pack.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.tbo_host_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def run_host(lib, bases, offsets, lo, hi, flags, p):
    bases = np.ascontiguousarray(bases, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    lo = np.ascontiguousarray(lo, np.int32)
    hi2 = np.array(hi, np.int32, copy=True)
    flags = np.ascontiguousarray(flags, np.uint8)
    n = len(offsets) - 1
    ins = np.full(n // 2, -9, np.int32)
    amb = np.zeros(n // 2, np.uint8)
    st = np.zeros(2, np.int64)
    paths = np.zeros(4, np.int64)
    lib.tbo_host_process(bases.ctypes.data, offsets.ctypes.data, n, lo.ctypes.data, hi2.ctypes.data, flags.ctypes.data,
                         p.min_overlap0, p.min_overlap, p.min_insert0, p.min_insert, p.max_ratio, p.min_second_ratio,
                         p.ratio_margin, p.ratio_offset, COMP.ctypes.data, ins.ctypes.data, amb.ctypes.data, st.ctypes.data,
                         paths.ctypes.data)
    return hi2, ins, amb, st, paths


def planes_of(seq, W):
    """bit planes of a byte string, base i = bit 31-(i&31) of word i>>5"""
    h = np.zeros(W, np.uint32)
    l = np.zeros(W, np.uint32)
    n = np.zeros(W, np.uint32)
    code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
    for i, ch in enumerate(seq):
        bit = np.uint32(1 << (31 - (i & 31)))
        if ch == ord("N"):
            n[i >> 5] |= bit
        elif ch in code:
            if code[ch] & 1:
                l[i >> 5] |= bit
            if code[ch] & 2:
                h[i >> 5] |= bit
    return h, l, n


def test_cap_line_never_undercuts_the_exact_cap(host):
    """the inner loop's linear bound on the admissible mismatch count vs cap_of, for every overlap length"""
    host.tbo_host_check_cap_line.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(1)
    for ratio in [0.1001, 0.0501, 0.1, 0.05, 1e-4, 0.0135] + list(rng.random(200) * 0.11):
        for margin in (5.0, 9.0):
            assert host.tbo_host_check_cap_line(float(ratio), margin) == 0, (ratio, margin)


def test_ratio_cap_line_admits_every_count_that_can_lower_the_best_ratio(host):
    """findBestRatio's screen: a count c >= 1 only matters if (T[c] + offset) / ov < bestRatio0"""
    host.tbo_host_check_ratio_cap_line.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(2)
    for ratio in [0.1001, 0.0501, 0.1, 0.05, 1e-4, 0.0135, 0.3, 0.45, 0.9, 1.5, 40.0] + list(rng.random(100) * 0.2):
        for offset in (0.4, 0.5, 0.0, 0.55, -0.3):
            assert host.tbo_host_check_ratio_cap_line(float(ratio), offset) == 0, (ratio, offset)


@pytest.mark.parametrize("reverse", [0, 1])
def test_packing_matches_a_bytewise_restatement(host, reverse):
    rng = np.random.default_rng(5 + reverse)
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTN", np.uint8)
    for trial in range(300):
        length = int(rng.integers(0, 200)) if trial else 0
        start = int(rng.integers(0, 40))
        buf = rng.integers(0, 256, start + length + int(rng.integers(0, 40)) + 1).astype(np.uint8)  # junk around the read
        seq = alphabet[rng.integers(0, len(alphabet), length)]
        odd = trial % 3 == 0 and length > 0
        if odd:
            seq[int(rng.integers(0, length))] = rng.choice(np.frombuffer(b"acgtnUuRYKM-.*@[`{", np.uint8))
        buf[start:start + length] = seq
        W = (max(length, 16) + 31) // 32 + 4
        out = np.zeros(3 * W, np.uint32)
        what = host.tbo_host_pack(buf.ctypes.data, start, length, reverse, W, out.ctypes.data)
        assert bool(what & 1) == odd
        assert bool(what & 2) == bool(np.any(seq == ord("N")))
        if odd:
            continue  # the planes are not used: such pairs take the byte-wise path
        want_seq = COMP[seq[::-1] & 127] if reverse else seq
        h, l, n = planes_of(want_seq, W)
        # under an N the code planes are unspecified (every use masks them with the N plane)
        keep = ~n
        assert np.array_equal(out[2 * W:], n)
        assert np.array_equal(out[:W] & keep, h & keep)
        assert np.array_equal(out[W:2 * W] & keep, l & keep)
        # nothing beyond the read
        total = np.concatenate([out[:W], out[W:2 * W], out[2 * W:]])
        for k in range(3):
            for w in range(W):
                valid = min(32, max(0, length - 32 * w))
                assert int(total[k * W + w]) & ((1 << (32 - valid)) - 1) == 0


@pytest.mark.parametrize("strict,seed", [(True, 21), (False, 22), (True, 23)])
def test_small_pairs_match_the_oracle(host, strict, seed):
    bases, quals, offsets, lo, hi, flags = small_pairs(1500, seed)
    p = otbo.default_params(strict)
    whi, wins, wamb, wst = otbo.process(bases, None, offsets, lo, hi, flags, p)
    ghi, gins, gamb, gst, paths = run_host(host, bases, offsets, lo, hi, flags, p)
    assert np.array_equal(gins, wins)
    assert np.array_equal(gamb, wamb)
    assert np.array_equal(ghi, whi)
    assert list(gst) == list(wst)
    assert paths[1] > 50 and paths[2] > 50 and paths[3] > 10  # second loop, N planes and the byte path were all exercised


def test_cfg2_pairs_and_long_ragged_reads_match_the_oracle(host):
    bases, offsets = synth.paired_adapter_reads(3000, seed=31)
    n = len(offsets) - 1
    L = np.diff(offsets).astype(np.int32)
    z, f = np.zeros(n, np.int32), np.zeros(n, np.uint8)
    for p in (otbo.default_params(True), otbo.default_params(False)):
        want = otbo.process(bases, None, offsets, z, L, f, p)
        got = run_host(host, bases, offsets, z, L, f, p)
        for g, w in zip(got[:4], (want[0], want[1], want[2], want[3])):
            assert np.array_equal(g, w)
        assert want[3][0] > 100
    # ragged: lengths 0..1008, inserts shorter and longer than the reads, unaligned starts
    rng = np.random.default_rng(8)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = []
    for _ in range(120):
        n1, n2 = (int(x) for x in rng.integers(0, 1009, 2))
        if rng.random() < 0.5:
            n1, n2 = int(rng.integers(0, 200)), int(rng.integers(0, 200))
        ins = int(rng.integers(1, max(2, n1 + n2)))
        frag = acgt[rng.integers(0, 4, ins)]
        r1 = np.concatenate([frag, acgt[rng.integers(0, 4, 1100)]])[:n1].copy()
        r2 = np.concatenate([COMP[frag[::-1]], acgt[rng.integers(0, 4, 1100)]])[:n2].copy()
        for r in (r1, r2):
            for _ in range(int(rng.integers(0, 1 + len(r) // 15))):
                r[int(rng.integers(0, len(r)))] = acgt[int(rng.integers(0, 4))]
            if len(r) and rng.random() < 0.3:
                r[int(rng.integers(0, len(r)))] = ord("N")
        seqs += [r1, r2]
    bases = np.concatenate(seqs).astype(np.uint8)
    offsets = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(s) for s in seqs], out=offsets[1:])
    L = np.diff(offsets).astype(np.int32)
    z, f = np.zeros(len(L), np.int32), np.zeros(len(L), np.uint8)
    p = otbo.default_params(False)
    want = otbo.process(bases, None, offsets, z, L, f, p)
    got = run_host(host, bases, offsets, z, L, f, p)
    for g, w in zip(got[:4], want):
        assert np.array_equal(g, w)
    assert want[3][0] > 20


def test_borderline_mismatch_rates_and_repeats_match_the_oracle(host):
    """overlaps whose mismatch ratio straddles maxRatio (the screen's cap is derived from it) and tandem repeats, where
    many alignments of one pair compete"""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = []
    for i in range(3000):
        n1, n2 = (int(x) for x in rng.integers(60, 260, 2))
        ins = int(rng.integers(20, n1 + n2))
        if i % 4 == 0:  # tandem repeat of a short unit, with a few point changes
            unit = acgt[rng.integers(0, 4, int(rng.integers(1, 9)))]
            frag = np.tile(unit, ins // len(unit) + 1)[:ins].copy()
        else:
            frag = acgt[rng.integers(0, 4, ins)]
        r1 = np.concatenate([frag, acgt[rng.integers(0, 4, 300)]])[:n1].copy()
        r2 = np.concatenate([COMP[frag[::-1]], acgt[rng.integers(0, 4, 300)]])[:n2].copy()
        rate = rng.random() * 0.2
        for r in (r1, r2):
            hit = rng.random(len(r)) < rate / 2
            r[hit] = acgt[rng.integers(0, 4, int(hit.sum()))]
        seqs += [r1, r2]
    bases = np.concatenate(seqs).astype(np.uint8)
    offsets = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum([len(x) for x in seqs], out=offsets[1:])
    L = np.diff(offsets).astype(np.int32)
    z, f = np.zeros(len(L), np.int32), np.zeros(len(L), np.uint8)
    for p in (otbo.default_params(True), otbo.default_params(False)):
        want = otbo.process(bases, None, offsets, z, L, f, p)
        got = run_host(host, bases, offsets, z, L, f, p)
        for g, w in zip(got[:4], want):
            assert np.array_equal(g, w)
        assert 200 < want[3][0] < 5000  # some pairs trimmed, some not


def test_coincident_ns_do_not_count_against_the_screen(host):
    """N against N is no mismatch for the reference; the screen reads an N as A in BOTH mates (r2' included), so a true
    overlap full of coincident Ns must survive it"""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = []
    for i in range(400):
        ins = int(rng.integers(60, 140))
        frag = acgt[rng.integers(0, 4, ins)].copy()
        frag[rng.random(ins) < (0.05 + 0.3 * rng.random())] = ord("N")  # the same fragment positions are N in both mates
        r1 = np.concatenate([frag, acgt[rng.integers(0, 4, 200)]])[:150].copy()
        r2 = np.concatenate([COMP[frag[::-1]], acgt[rng.integers(0, 4, 200)]])[:150].copy()
        seqs += [r1, r2]
    bases = np.concatenate(seqs).astype(np.uint8)
    offsets = np.arange(0, 150 * len(seqs) + 1, 150, dtype=np.int64)
    L = np.full(len(seqs), 150, np.int32)
    z, f = np.zeros(len(L), np.int32), np.zeros(len(L), np.uint8)
    for p in (otbo.default_params(True), otbo.default_params(False)):
        want = otbo.process(bases, None, offsets, z, L, f, p)
        got = run_host(host, bases, offsets, z, L, f, p)
        for g, w in zip(got[:4], want):
            assert np.array_equal(g, w)
        assert want[3][0] > 100
