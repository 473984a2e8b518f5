"""CPU tests of the native FASTQ feed (include/fastq_b200.h) against the plain-Python reader/writer of
bbtools_b200/fasta.py + bbduk.py on synthetic FASTQ text: CRLF, no final newline, empty reads, '@'/'+' inside
quality lines, two-file and interleaved pairs, many threads (records straddling every block boundary)."""
import numpy as np
import pytest

from bbtools_b200 import F_REMOVED
from bbtools_b200.fastq import FastqBatch


def make_fastq(n, seed, crlf=False, final_newline=True, max_len=120):
    rng = np.random.default_rng(seed)
    names, seqs, quals, lines = [], [], [], []
    eol = b"\r\n" if crlf else b"\n"
    for i in range(n):
        ln = int(rng.integers(0, max_len + 1))
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), ln))
        q = bytes(rng.integers(33, 75, ln).astype(np.uint8))  # includes '@' (64) and '+' (43)
        if ln and i % 7 == 0:
            q = b"@" + q[1:]
        name = b"read%d/%d some comment @+ here" % (i, seed)
        names.append(name)
        seqs.append(s)
        quals.append(q)
        lines += [b"@" + name, s, b"+" + (name if i % 5 == 0 else b""), q]
    text = eol.join(lines) + (eol if final_newline else b"")
    return np.frombuffer(text, np.uint8), names, seqs, quals


@pytest.mark.parametrize("crlf,final_nl,threads", [(False, True, 1), (False, True, 7), (True, True, 3), (False, False, 5),
                                                    (True, False, 16)])
def test_index_and_gather_match_python(crlf, final_nl, threads):
    text, names, seqs, quals = make_fastq(5000, 1, crlf, final_nl)
    fb = FastqBatch(text, threads=threads)
    assert fb.n_reads == 5000
    bases, offsets = fb.arrays()
    assert np.array_equal(np.diff(offsets), [len(s) for s in seqs])
    assert bytes(bases) == b"".join(seqs)
    assert bytes(fb.quals()) == b"".join(quals)
    rec = fb.rec.reshape(-1, 4)
    t = bytes(text)
    for i in (0, 1, 17, 2500, 4999):
        assert t[rec[i, 0]:rec[i, 0] + 1 + len(names[i])] == b"@" + names[i]
        assert t[rec[i, 3]:rec[i, 3] + rec[i, 2]] == quals[i]


def python_format(names, seqs, quals, per, lo, hi, flags, removed, mate_sel, trim_removed):
    out = []
    for u in range(len(seqs) // per):
        rem = bool(flags[u * per] & F_REMOVED)
        if rem != removed:
            continue
        for q in range(per):
            if mate_sel and not ((per == 2 and mate_sel == q + 1) or (per == 1 and mate_sel == 1)):
                continue
            i = u * per + q
            a, b = (int(lo[i]), int(hi[i])) if (not rem or trim_removed) else (0, len(seqs[i]))
            out.append(b"@" + names[i] + b"\n" + seqs[i][a:b] + b"\n+\n" + quals[i][a:b] + b"\n")
    return b"".join(out)


def test_two_files_interleave_and_format():
    t1, n1, s1, q1 = make_fastq(3001, 2)
    t2, n2, s2, q2 = make_fastq(3001, 3, crlf=True)
    fb = FastqBatch(t1, t2, threads=6)
    names = [x for p in zip(n1, n2) for x in p]
    seqs = [x for p in zip(s1, s2) for x in p]
    quals = [x for p in zip(q1, q2) for x in p]
    bases, offsets = fb.arrays()
    assert bytes(bases) == b"".join(seqs)
    rng = np.random.default_rng(9)
    L = np.diff(offsets)
    lo = np.minimum(rng.integers(0, 5, len(L)), L).astype(np.int32)
    hi = np.maximum(lo, L - rng.integers(0, 40, len(L))).astype(np.int32)
    flags = np.zeros(len(L), np.uint8)
    rem = rng.random(len(L) // 2) < 0.2
    flags[0::2][rem] = F_REMOVED
    flags[1::2][rem] = F_REMOVED
    for removed, mate_sel, ottm in ((False, 0, False), (False, 1, False), (False, 2, False), (True, 0, False), (True, 0, True),
                                    (True, 2, False)):
        got = bytes(fb.format(2, lo, hi, flags, removed=removed, mate_sel=mate_sel, trim_removed=ottm))
        assert got == python_format(names, seqs, quals, 2, lo, hi, flags, removed, mate_sel, ottm), (removed, mate_sel, ottm)
    # single-end view of the first file
    fb1 = FastqBatch(t1, threads=4)
    L1 = np.array([len(s) for s in s1])
    z = np.zeros(len(L1), np.int32)
    got = bytes(fb1.format(1, z, L1.astype(np.int32), np.zeros(len(L1), np.uint8)))
    assert got == python_format(n1, s1, q1, 1, z, L1, np.zeros(len(L1), np.uint8), False, 0, False)


def test_malformed_and_empty():
    assert FastqBatch(np.zeros(0, np.uint8)).n_reads == 0
    with pytest.raises(ValueError):
        FastqBatch(np.frombuffer(b"@r\nACGT\n+\nIII\n", np.uint8))  # quality length differs
    with pytest.raises(ValueError):
        FastqBatch(np.frombuffer(b"r\nACGT\n+\nIIII\n", np.uint8))  # no '@'
    with pytest.raises(ValueError):
        FastqBatch(np.frombuffer(b"@r\nACGT\n-\nIIII\n", np.uint8))  # no '+'


@pytest.mark.parametrize("block,gz,final_nl", [(997, False, True), (4096, True, False), (50, False, False), (1 << 20, False, True)])
def test_bounded_streaming_of_mate_files(tmp_path, block, gz, final_nl):
    """iter_fastq_blocks: whole records only, the same number from both files whatever their byte sizes, nothing lost"""
    import gzip

    from bbtools_b200.fastq import iter_fastq_blocks
    t1, n1, s1, q1 = make_fastq(1200, 11, final_newline=final_nl)
    t2, n2, s2, q2 = make_fastq(1200, 12, crlf=True, final_newline=final_nl, max_len=40)  # much shorter records
    p1, p2 = tmp_path / ("a.fq.gz" if gz else "a.fq"), tmp_path / "b.fq"
    with (gzip.open(p1, "wb") if gz else open(p1, "wb")) as f:
        f.write(t1.tobytes())
    with open(p2, "wb") as f:
        f.write(t2.tobytes())
    seqs = []
    got1, got2 = [], []
    for x1, x2 in iter_fastq_blocks(str(p1), str(p2), block_bytes=block):
        fb = FastqBatch(x1, x2)
        assert fb.n_reads > 0 and fb.n_reads % 2 == 0
        b, off = fb.arrays()
        seqs += [b[off[i]:off[i + 1]].tobytes() for i in range(fb.n_reads)]
        got1.append(x1.tobytes())
        got2.append(x2.tobytes())
    assert seqs[0::2] == s1 and seqs[1::2] == s2
    assert b"".join(got1) == t1.tobytes() and b"".join(got2) == t2.tobytes()
    # single interleaved file: blocks hold whole pairs
    n_tot = 0
    for x1, x2 in iter_fastq_blocks(str(p2), None, block_bytes=block, unit=2):
        assert x2 is None
        n = FastqBatch(x1).n_reads
        assert n % 2 == 0
        n_tot += n
    assert n_tot == 1200


def test_bounded_streaming_reports_truncated_and_unequal_inputs(tmp_path):
    from bbtools_b200.fastq import iter_fastq_blocks
    t1, *_ = make_fastq(300, 21)
    t2, *_ = make_fastq(299, 22)
    p1, p2, p3 = tmp_path / "a.fq", tmp_path / "b.fq", tmp_path / "c.fq"
    open(p1, "wb").write(t1.tobytes())
    open(p2, "wb").write(t2.tobytes())
    open(p3, "wb").write(t1.tobytes()[:-200] if t1.tobytes()[-201:-200] != b"\n" else t1.tobytes()[:-199])
    with pytest.raises(ValueError):
        list(iter_fastq_blocks(str(p1), str(p2), block_bytes=2000))  # 300 vs 299 records
    with pytest.raises(ValueError):
        list(iter_fastq_blocks(str(p3), None, block_bytes=2000))  # ends inside a record
    assert list(iter_fastq_blocks(str(tmp_path / "a.fq"), None, block_bytes=10 ** 9))[0][0].size == t1.size
    open(tmp_path / "e.fq", "wb").close()
    assert list(iter_fastq_blocks(str(tmp_path / "e.fq"), None, block_bytes=100)) == []
