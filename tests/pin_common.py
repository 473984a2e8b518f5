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
