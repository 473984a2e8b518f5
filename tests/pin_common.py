"""Shared by tools/pin_reference.py (which runs the REAL reference where a JRE exists and records digests of what it wrote)
and tests/test_reference_pin.py (which holds the oracle and the GPU tool to those digests): the pinned command lines, the
seeded inputs, and the rendering of (lo, hi, flags) into the FASTQ bytes the reference writes with ordered=t.

Every case is a bbduk.sh command line in the reference's own syntax (jgi/BBDuk.java:120-700 parser), so the same strings
go to `java jgi.BBDuk`, `java bbduk.BBDukS` and bbtools_b200.bbduk.parse_args."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIGESTS = os.path.join(GOLDEN, "reference_digests.json")

# name -> (input set, flags). Inputs: "pairs" = 20 k cfg-2 pairs (seed 1), "single" = the same reads as 40 k single reads,
# "ragged" = 6 k ragged single reads with N / IUPAC / lower case and adapter fragments (seed 4).
CASES = {
    "cfg1_ktrim_r_k23": ("single", ["ktrim=r", "k=23"]),
    "cfg2_ktrim_r_k23_mink11_hdist1_tpe": ("pairs", ["ktrim=r", "k=23", "mink=11", "hdist=1", "tpe"]),
    "ktrim_l_mink11_hdist1": ("pairs", ["ktrim=l", "k=23", "mink=11", "hdist=1"]),
    "ktrim_r_ragged": ("ragged", ["ktrim=r", "k=23", "mink=11", "hdist=1", "minlen=1"]),
    "ktrim_r_exclusive_pad": ("pairs", ["ktrim=r", "k=23", "mink=8", "hdist=1", "hdist2=0", "ktrimexclusive=t", "trimpad=2"]),
    "ktrim_r_restrict": ("pairs", ["ktrim=r", "k=27", "hdist=1", "restrictright=50"]),
    "ktrim_r_forbidn_mlf": ("ragged", ["ktrim=r", "k=21", "hdist=0", "forbidn=t", "mm=f", "mlf=0.5"]),
    "kfilter_k31": ("pairs", ["k=31"]),
    "kfilter_k31_mbk2": ("pairs", ["k=31", "mbk=2", "mm=f"]),
    "kfilter_mkf": ("pairs", ["k=27", "hdist=1", "mkf=0.05"]),
    "kfilter_mcf": ("pairs", ["k=25", "mcf=0.2"]),
    "kfilter_k40": ("pairs", ["k=40"]),
    "kfilter_qhdist1": ("single", ["k=23", "qhdist=1"]),
    "kfilter_speed5": ("single", ["k=20", "speed=5"]),
    "kfilter_qskip3": ("single", ["k=20", "qskip=3"]),
    "kfilter_rcomp_f": ("ragged", ["k=19", "rcomp=f", "mm=3"]),
    "kmask_N": ("single", ["ktrim=N", "k=23", "mink=11", "hdist=1"]),
}
MAIN_CLASSES = ("jgi.BBDuk", "bbduk.BBDukS")  # bbdukOld.sh:388 and bbduk.sh:391


def inputs(kind):
    """-> (bases, offsets, paired) of one input set"""
    from bbtools_b200 import synth
    if kind in ("pairs", "single"):
        b, off = synth.paired_adapter_reads(20000, seed=1)
        return b, off, kind == "pairs"
    seqs = [ln.strip() for ln in open(os.path.join(GOLDEN, "adapters.fa")) if not ln.startswith(">")]
    b, off = synth.ragged_reads(6000, seed=4, adapter=seqs[0].encode())
    return b, off, False


def record(i, paired, seq, qual=None):
    name = b"@r%d %d:N:0" % ((i // 2, 1 + (i & 1)) if paired else (i, 1))
    return name + b"\n" + seq + b"\n+\n" + (qual if qual is not None else b"I" * len(seq)) + b"\n"


def write_inputs(kind, directory):
    """writes the FASTQ file(s) of one input set, returns their paths"""
    b, off, paired = inputs(kind)
    paths = [os.path.join(directory, f"{kind}_{m}.fq") for m in ((1, 2) if paired else (1,))]
    files = [open(p, "wb") for p in paths]
    for i in range(len(off) - 1):
        files[i & 1 if paired else 0].write(record(i, paired, bytes(b[off[i]:off[i + 1]])))
    for f in files:
        f.close()
    return paths


def render(bases, offsets, paired, lo, hi, flags, maskbits=None, mask_off=None):
    """the bytes of out / out2 / outm / outm2 as the reference writes them with ordered=t: kept reads cut to [lo, hi)
    (kmask: covered bases as N, qualities 0), removed pairs whole and untrimmed into outm (jgi/BBDuk.java:2580-2700)"""
    from bbtools_b200 import F_REMOVED
    out = [[], [], [], []]
    for i in range(len(offsets) - 1):
        unit = i & ~1 if paired else i
        rem = bool(flags[unit] & F_REMOVED)
        s = bytearray(bases[offsets[i]:offsets[i + 1]])
        q = bytearray(b"I" * len(s))
        if maskbits is not None:
            w0 = int(mask_off[i])
            for j in range(len(s)):
                if (int(maskbits[w0 + (j >> 5)]) >> (j & 31)) & 1:
                    s[j] = ord("N")
                    q[j] = ord("!")
        a, b_ = (0, len(s)) if rem else (int(lo[i]), int(hi[i]))
        out[(2 if rem else 0) + ((i & 1) if paired else 0)].append(record(i, paired, bytes(s[a:b_]), bytes(q[a:b_])))
    return [b"".join(x) for x in out]


def sha(data):
    return hashlib.sha256(data).hexdigest()


# ---- Seal (jgi.Seal, seal.sh): name -> flags; inputs: seal_refs() / seal_reads() below, always paired ---------------------
SEAL_CASES = {
    "seal_default": [],
    "seal_ambig_all_cz3": ["ambig=all", "cz=3"],
    "seal_ambig_first_k25_mm_f": ["ambig=first", "k=25", "mm=f"],
    "seal_ambig_toss_mkh5": ["ambig=toss", "mkh=5"],
    "seal_kpt_f": ["kpt=f", "ambig=all"],
    "seal_hdist1_k21": ["k=21", "hdist=1"],
    "seal_match_unique": ["match=unique", "ambig=all"],
    "seal_match_first_qskip3": ["match=first", "qskip=3"],
    "seal_czf_mkf": ["czf=0.05", "mkf=0.2", "ambig=all"],
    "seal_restrict_speed": ["restrictleft=80", "speed=4", "k=27"],
}
SEAL_CLASS = "jgi.Seal"  # seal.sh


def seal_refs():
    """8 seeded reference sequences of 600 bases: 1/2 and 5/6 are strains 2 % apart, 3 carries an N and an IUPAC code"""
    rng = np.random.default_rng(21)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    refs = [acgt[rng.integers(0, 4, 600)].copy() for _ in range(8)]
    for a, b in ((0, 1), (4, 5)):
        refs[b] = refs[a].copy()
        q = rng.integers(0, 600, 12)
        refs[b][q] = acgt[rng.integers(0, 4, 12)]
    refs[2][100] = ord("N")
    refs[2][300] = ord("R")
    return [(f"seq{i + 1}", bytes(r)) for i, r in enumerate(refs)]


def seal_reads(n_pairs=4000):
    """-> (bases, offsets) of n_pairs seeded 2 x 120 bp pairs: both mates from one reference (either strand), 1 % substitutions,
    0.2 % N, every tenth pair random"""
    rng = np.random.default_rng(22)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    refs = [np.frombuffer(s, np.uint8) for _, s in seal_refs()]
    comp = np.zeros(256, np.uint8)
    for a, c in zip(b"ACGTNR", b"TGCANY"):
        comp[a] = c
    rows = []
    for p in range(n_pairs):
        src = refs[int(rng.integers(0, 8))]
        for _ in range(2):
            if p % 10 == 9:
                r = acgt[rng.integers(0, 4, 120)]
            else:
                s = int(rng.integers(0, 480))
                r = src[s:s + 120].copy()
                e = rng.random(120)
                r = np.where(e < 0.01, acgt[rng.integers(0, 4, 120)], r)
                r = np.where(e > 0.998, np.uint8(ord("N")), r)
                if rng.integers(0, 2):
                    r = comp[r][::-1]
            rows.append(r.astype(np.uint8))
    bases = np.concatenate(rows)
    return bases, np.arange(2 * n_pairs + 1, dtype=np.int64) * 120


def write_seal_inputs(directory):
    ref = os.path.join(directory, "seal_refs.fa")
    with open(ref, "wb") as f:
        for name, s in seal_refs():
            f.write(b">" + name.encode() + b"\n" + s + b"\n")
    b, off = seal_reads()
    paths = [os.path.join(directory, f"seal_{m}.fq") for m in (1, 2)]
    files = [open(p, "wb") for p in paths]
    for i in range(len(off) - 1):
        files[i & 1].write(record(i, True, bytes(b[off[i]:off[i + 1]])))
    for f in files:
        f.close()
    return ref, paths


def render_seal(bases, offsets, hit):
    """outm / outm2 / outu / outu2 as seal.sh writes them with ordered=t: whole pairs, untrimmed (jgi/Seal.java:2278-2286)"""
    out = [[], [], [], []]
    for i in range(len(offsets) - 1):
        out[(0 if hit[i // 2] else 2) + (i & 1)].append(record(i, True, bytes(bases[offsets[i]:offsets[i + 1]])))
    return [b"".join(x) for x in out]
