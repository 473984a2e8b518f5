"""Counter-based synthetic inputs for the BBDuk configs of BASELINE.json.

Every byte is a pure function of (seed, read index, position) through splitmix64, so any chunk of a
workload can be regenerated independently -- on the host with numpy (here) or on the device by
csrc/synth.cu, which implements the same formulas and is tested bit-for-bit against this file.
No reference code is involved (the reference's generators, synth.RandomReads3, need a JVM).
"""
import numpy as np

# TruSeq adapter prefixes named in SURVEY.md 8d (both occur inside tests/golden/adapters.fa records)
ADAPTER_R1 = b"AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"
ADAPTER_R2 = b"AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_C2 = np.uint64(0xD1B54A32D192ED03)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_ACGT = np.frombuffer(b"ACGT", np.uint8)


def mix64(z):
    """splitmix64 finalizer on a uint64 array (wrapping arithmetic)."""
    z = np.asarray(z, np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def rnd(seed, stream, idx):
    """rnd(seed, stream, idx) = mix64(seed + stream*GOLD + idx*C2)  (all mod 2^64)"""
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed) + np.uint64(stream) * _GOLD + np.asarray(idx, np.uint64) * _C2)


def _comp_ascii(b):
    out = b.copy()
    out[b == ord("A")] = ord("T")
    out[b == ord("C")] = ord("G")
    out[b == ord("G")] = ord("C")
    out[b == ord("T")] = ord("A")
    return out


def _apply_errors(reads, seed, stream, gidx, sub_per_10k, n_per_10k):
    """substitutions then N's; gidx = global flat base index (uint64, same shape as reads)."""
    e = rnd(seed, stream, gidx)
    code = (((reads >> 1) ^ (reads >> 2)) & 3).astype(np.uint64)  # A0 C1 G2 T3
    sub = (e % np.uint64(10000)) < np.uint64(sub_per_10k)
    newcode = (code + np.uint64(1) + ((e >> np.uint64(16)) % np.uint64(3))) & np.uint64(3)
    reads = np.where(sub, _ACGT[newcode.astype(np.int64)], reads)
    isn = ((e >> np.uint64(32)) % np.uint64(10000)) < np.uint64(n_per_10k)
    return np.where(isn, np.uint8(ord("N")), reads).astype(np.uint8)


def insert_sizes(seed, pair_idx):
    """cfg-2 insert-size mixture: 70 % U[300,600], 25 % U[35,149], 5 % U[0,34]."""
    r = rnd(seed, 0, pair_idx)
    cls = r % np.uint64(100)
    v = r >> np.uint64(8)
    ins = np.where(cls < 70, np.uint64(300) + v % np.uint64(301),
                   np.where(cls < 95, np.uint64(35) + v % np.uint64(115), v % np.uint64(35)))
    return ins.astype(np.int64)


def paired_adapter_reads(n_pairs, first_pair=0, read_len=150, seed=1, sub_per_10k=50, n_per_10k=5):
    """cfg 2 / cfg 4 reads: interleaved 2 x read_len bp pairs with adapter read-through on short inserts.

    Returns (bases uint8[2*n_pairs*read_len], offsets int64[2*n_pairs+1]); read 2i = R1, 2i+1 = R2."""
    L = read_len
    p = np.arange(first_pair, first_pair + n_pairs, dtype=np.uint64)
    j = np.arange(L, dtype=np.uint64)
    gidx = p[:, None] * np.uint64(L) + j[None, :]
    r1 = _ACGT[(rnd(seed, 1, gidx) & np.uint64(3)).astype(np.int64)]
    r2 = _ACGT[(rnd(seed, 2, gidx) & np.uint64(3)).astype(np.int64)]
    ins = insert_sizes(seed, p)
    jj = np.arange(L, dtype=np.int64)[None, :]
    I = ins[:, None]
    short = I < L
    # R2[j] = comp(R1[I-1-j]) for j < I
    src = np.clip(I - 1 - jj, 0, L - 1)
    rc = _comp_ascii(np.take_along_axis(r1, src, axis=1))
    r2 = np.where(short & (jj < I), rc, r2)
    a1 = np.frombuffer(ADAPTER_R1, np.uint8)
    a2 = np.frombuffer(ADAPTER_R2, np.uint8)
    ai = jj - I
    in_ad = short & (ai >= 0) & (ai < len(a1))
    aic = np.clip(ai, 0, len(a1) - 1)
    r1 = np.where(in_ad, a1[aic], r1)
    r2 = np.where(in_ad, a2[aic], r2)
    r1 = _apply_errors(r1, seed, 3, gidx, sub_per_10k, n_per_10k)
    r2 = _apply_errors(r2, seed, 4, gidx, sub_per_10k, n_per_10k)
    out = np.empty((n_pairs, 2, L), np.uint8)
    out[:, 0, :] = r1
    out[:, 1, :] = r2
    offsets = np.arange(0, (2 * n_pairs + 1) * L, L, dtype=np.int64)
    return out.reshape(-1), offsets


def single_adapter_reads(n_reads, adapter, first_read=0, read_len=150, seed=1, frac_pct=30, min_off=20,
                         sub_per_10k=50, n_per_10k=5):
    """cfg 1 reads: SE, uniform ACGT; frac_pct % carry `adapter` spliced at a uniform offset in
    [min_off, read_len-1] followed by the random tail."""
    L = read_len
    r = np.arange(first_read, first_read + n_reads, dtype=np.uint64)
    j = np.arange(L, dtype=np.uint64)
    gidx = r[:, None] * np.uint64(L) + j[None, :]
    reads = _ACGT[(rnd(seed, 1, gidx) & np.uint64(3)).astype(np.int64)]
    c = rnd(seed, 0, r)
    has = (c % np.uint64(100)) < np.uint64(frac_pct)
    off = (np.uint64(min_off) + (c >> np.uint64(8)) % np.uint64(L - min_off)).astype(np.int64)
    ad = np.frombuffer(bytes(adapter), np.uint8)
    jj = np.arange(L, dtype=np.int64)[None, :]
    ai = jj - off[:, None]
    in_ad = has[:, None] & (ai >= 0) & (ai < len(ad))
    reads = np.where(in_ad, ad[np.clip(ai, 0, len(ad) - 1)], reads)
    reads = _apply_errors(reads, seed, 3, gidx, sub_per_10k, n_per_10k)
    offsets = np.arange(0, (n_reads + 1) * L, L, dtype=np.int64)
    return reads.reshape(-1), offsets


def random_reference(n_scaffolds, scaffold_len, seed=7):
    """cfg 3 / cfg 4 reference: n scaffolds of uniform ACGT."""
    total = n_scaffolds * scaffold_len
    out = np.empty(total, np.uint8)
    step = 1 << 22
    for s in range(0, total, step):
        idx = np.arange(s, min(total, s + step), dtype=np.uint64)
        out[s:s + len(idx)] = _ACGT[(rnd(seed, 5, idx) & np.uint64(3)).astype(np.int64)]
    offsets = np.arange(0, total + 1, scaffold_len, dtype=np.int64)
    return out, offsets


def contaminant_reads(n_reads, ref_bases, first_read=0, read_len=150, seed=1, contam_pct=10, sub_per_10k=100,
                      n_per_10k=5):
    """cfg 3 reads: contam_pct % drawn from the reference (forward strand, uniform start over the
    concatenated reference, 1 % substitutions), the rest uniform ACGT."""
    L = read_len
    r = np.arange(first_read, first_read + n_reads, dtype=np.uint64)
    j = np.arange(L, dtype=np.uint64)
    gidx = r[:, None] * np.uint64(L) + j[None, :]
    reads = _ACGT[(rnd(seed, 1, gidx) & np.uint64(3)).astype(np.int64)]
    c = rnd(seed, 0, r)
    has = (c % np.uint64(100)) < np.uint64(contam_pct)
    span = np.uint64(len(ref_bases) - L + 1)
    start = ((c >> np.uint64(8)) % span).astype(np.int64)
    src = start[:, None] + np.arange(L, dtype=np.int64)[None, :]
    reads = np.where(has[:, None], np.asarray(ref_bases)[src], reads)
    reads = _apply_errors(reads, seed, 3, gidx, sub_per_10k, n_per_10k)
    offsets = np.arange(0, (n_reads + 1) * L, L, dtype=np.int64)
    return reads.reshape(-1), offsets


def ragged_reads(n_reads, seed=3, min_len=0, max_len=400, alphabet=b"ACGTNacgtnRYKMUu", adapter=None):
    """Edge-case reads for parity tests: ragged lengths (including 0), lowercase, N / IUPAC, optional
    adapter fragments at random places."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, n_reads)
    alpha = np.frombuffer(alphabet, np.uint8)
    # bias towards plain ACGT so k-mers survive
    w = np.ones(len(alpha))
    w[:4] = 40.0
    w /= w.sum()
    seqs = []
    for n in lens:
        s = alpha[rng.choice(len(alpha), size=int(n), p=w)]
        if adapter is not None and n > 8 and rng.random() < 0.6:
            a = np.frombuffer(bytes(adapter), np.uint8)
            frag_len = int(rng.integers(5, len(a) + 1))
            fs = int(rng.integers(0, len(a) - frag_len + 1))
            pos = int(rng.integers(0, n))
            frag = a[fs:fs + frag_len][: n - pos]
            if rng.random() < 0.3:  # reverse-complement strand
                frag = _comp_ascii(frag[::-1].copy())
            s[pos:pos + len(frag)] = frag
        seqs.append(s)
    offsets = np.zeros(n_reads + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    bases = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
    return bases.astype(np.uint8), offsets


def genome_bases(genome_len, seed=11, start=0, n=None):
    """cfg 5 genome: base i = ACGT[rnd(seed, 5, i) & 3] for i in [start, start+n)."""
    n = genome_len - start if n is None else n
    out = np.empty(n, np.uint8)
    step = 1 << 22
    for s in range(0, n, step):
        idx = np.arange(start + s, start + min(n, s + step), dtype=np.uint64)
        out[s:s + len(idx)] = _ACGT[(rnd(seed, 5, idx) & np.uint64(3)).astype(np.int64)]
    return out


def genome_reads(n_reads, genome_len, first_read=0, read_len=150, seed=11, sub_per_10k=10):
    """cfg 5 reads: SE, sampled uniformly from the synthetic genome (either strand), substitutions only.
    Same formulas as csrc/kcount.cu:kc_synth_kernel."""
    L = read_len
    r = np.arange(first_read, first_read + n_reads, dtype=np.uint64)
    c = rnd(seed, 0, r)
    rev = (c & np.uint64(1)) != 0
    start = ((c >> np.uint64(8)) % np.uint64(genome_len - L + 1)).astype(np.int64)
    j = np.arange(L, dtype=np.int64)[None, :]
    gi = start[:, None] + np.where(rev[:, None], L - 1 - j, j)
    code = (rnd(seed, 5, gi.astype(np.uint64)) & np.uint64(3)).astype(np.int64)
    code = np.where(rev[:, None], 3 - code, code)
    e = rnd(seed, 3, r[:, None] * np.uint64(L) + j.astype(np.uint64))
    sub = (e % np.uint64(10000)) < np.uint64(sub_per_10k)
    newcode = (code + 1 + ((e >> np.uint64(16)) % np.uint64(3)).astype(np.int64)) & 3
    code = np.where(sub, newcode, code)
    offsets = np.arange(0, (n_reads + 1) * L, L, dtype=np.int64)
    return _ACGT[code].astype(np.uint8).reshape(-1), offsets
