"""Seal's k-mer matching path on B200 (include/seal_b200.h): ctypes mirror of the POD structs and the host-side class.

Mirrors jgi.Seal's loader + matching block (jgi/Seal.java:1760-1946, :2186-2276). No CPU fallback: the CUDA library
must be built (bbtools_b200/_lib.py raises ImportError otherwise)."""
import ctypes as C

import numpy as np

AMBIG_ALL, AMBIG_FIRST, AMBIG_TOSS, AMBIG_RANDOM = 1, 2, 3, 4
MATCH_ALL, MATCH_FIRST, MATCH_UNIQUE = 1, 2, 3


class SealCfg(C.Structure):
    """include/seal_b200.h seal_cfg (keep the field order in sync with the header)."""
    _fields_ = [
        ("struct_size", C.c_int32),
        ("k", C.c_int32),
        ("rcomp", C.c_int32),
        ("mask_middle", C.c_int32),
        ("mid_mask_len", C.c_int32),
        ("forbid_ns", C.c_int32),
        ("hdist", C.c_int32),
        ("speed", C.c_int32),
        ("qskip", C.c_int32),
        ("rskip", C.c_int32),
        ("restrict_left", C.c_int32),
        ("restrict_right", C.c_int32),
        ("ambig_mode", C.c_int32),
        ("match_mode", C.c_int32),
        ("keep_pairs_together", C.c_int32),
        ("clearzone", C.c_int32),
        ("clearzone_fraction", C.c_float),
        ("min_kmer_hits", C.c_int32),
        ("min_kmer_fraction", C.c_float),
        ("device", C.c_int32),
        ("table_load_pct", C.c_int32),
        ("ids_stride", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


class SealOut(C.Structure):
    _fields_ = [("n_assigned", C.c_void_p), ("first_id", C.c_void_p), ("n_sites", C.c_void_p), ("max_hits", C.c_void_p),
                ("ids", C.c_void_p)]


class SealStats(C.Structure):
    _fields_ = [("reads_in", C.c_int64), ("bases_in", C.c_int64), ("reads_matched", C.c_int64), ("bases_matched", C.c_int64),
                ("reads_unmatched", C.c_int64), ("bases_unmatched", C.c_int64), ("reserved", C.c_int64 * 2)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


def make_cfg(**kw) -> SealCfg:
    """Seal's defaults (jgi/Seal.java:104-128, :3088-3098) with overrides."""
    c = SealCfg()
    c.struct_size = C.sizeof(SealCfg)
    c.k, c.rcomp, c.mask_middle, c.mid_mask_len, c.forbid_ns, c.hdist = 31, 1, 1, 0, 0, 0
    c.ambig_mode, c.match_mode, c.keep_pairs_together = AMBIG_RANDOM, MATCH_ALL, 1
    c.min_kmer_hits, c.table_load_pct, c.ids_stride = 1, 50, 4
    for k, v in kw.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


class SealResult:
    def __init__(self, n_units, stride):
        self.n_assigned = np.zeros(n_units, np.int32)
        self.first_id = np.zeros(n_units, np.int32)
        self.n_sites = np.zeros(n_units, np.int32)
        self.max_hits = np.zeros(n_units, np.int32)
        self.ids = np.zeros(max(1, n_units * max(stride, 0)), np.int32)
        self.stride = stride

    def struct(self):
        o = SealOut()
        o.n_assigned, o.first_id = self.n_assigned.ctypes.data, self.first_id.ctypes.data
        o.n_sites, o.max_hits = self.n_sites.ctypes.data, self.max_hits.ctypes.data
        o.ids = self.ids.ctypes.data if self.stride > 0 else None
        return o

    def fields(self):
        return {"n_assigned": self.n_assigned, "first_id": self.first_id, "n_sites": self.n_sites, "max_hits": self.max_hits,
                "ids": self.ids}


def n_units(cfg, n_reads, paired):
    return n_reads // 2 if (paired and cfg.keep_pairs_together) else n_reads



class SealIndexGPU:
    """Reference table + matcher on one GPU: add_ref -> finalize -> process (host numpy buffers)."""

    def __init__(self, cfg: SealCfg):
        from . import _lib
        self.lib = _lib.load()
        self.cfg = cfg
        h = C.c_void_p()
        if self.lib.seal_b200_create(C.byref(cfg), C.byref(h)):
            raise ValueError(self.lib.seal_b200_last_error(None).decode())
        self.h = h
        self.n_seqs = 0

    def _check(self, rc):
        if rc:
            raise RuntimeError(self.lib.seal_b200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.seal_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_ref(self, bases, offsets):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        self._check(self.lib.seal_b200_add_ref(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1))
        self.n_seqs += len(offsets) - 1

    def finalize(self):
        """(storedKmers, (k-mer, id) entries, refKmers)"""
        v = np.zeros(3, np.int64)
        self._check(self.lib.seal_b200_finalize(self.h, v.ctypes.data))
        return tuple(int(x) for x in v)

    def table(self):
        n = C.c_int64()
        self._check(self.lib.seal_b200_table_export(self.h, None, None, 0, C.byref(n)))
        keys, ids = np.zeros(n.value, np.uint64), np.zeros(n.value, np.int32)
        self._check(self.lib.seal_b200_table_export(self.h, keys.ctypes.data, ids.ctypes.data, n.value, C.byref(n)))
        return keys, ids

    def process(self, bases, offsets, paired, first_numeric_id=0):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        nu = self.lib.seal_b200_n_units(self.h, n, 1 if paired else 0)
        res = SealResult(nu, self.cfg.ids_stride)
        st = SealStats()
        o = res.struct()
        self._check(self.lib.seal_b200_process(self.h, bases.ctypes.data, offsets.ctypes.data, n, 1 if paired else 0,
                                               first_numeric_id, C.byref(o), C.byref(st)))
        return res, st

    def scaffold_counts(self):
        n = self.n_seqs + 1
        a = [np.zeros(n, np.int64) for _ in range(4)]
        self._check(self.lib.seal_b200_scaffold_counts(self.h, *(x.ctypes.data for x in a), n))
        return a

    @property
    def launches(self):
        return int(self.lib.seal_b200_launch_count(self.h))


# ---- seal.sh's flag surface (jgi/Seal.java:150-380) -------------------------------------------------------------------
def _parse_bool(b):
    """shared/Parse.java parseBoolean: a bare flag is true; t/true/1 and f/false/0."""
    if b is None:
        return True
    x = b.lower()
    if x in ("t", "true", "1"):
        return True
    if x in ("f", "false", "0"):
        return False
    raise ValueError(f"not a boolean: {b}")


_AMBIG = {"keep": AMBIG_FIRST, "best": AMBIG_FIRST, "first": AMBIG_FIRST, "all": AMBIG_ALL, "random": AMBIG_RANDOM, "rand": AMBIG_RANDOM,
          "toss": AMBIG_TOSS, "discard": AMBIG_TOSS, "remove": AMBIG_TOSS}
_MATCH = {"all": MATCH_ALL, "best": MATCH_ALL, "first": MATCH_FIRST, "unique": MATCH_UNIQUE, "firstunique": MATCH_UNIQUE}
_INT_FLAGS = {"k": "k", "hdist": "hdist", "hammingdistance": "hdist", "skip": "rskip", "refskip": "rskip", "rskip": "rskip",
              "qskip": "qskip", "speed": "speed", "minkmerhits": "min_kmer_hits", "minhits": "min_kmer_hits", "mh": "min_kmer_hits",
              "mkh": "min_kmer_hits", "restrictleft": "restrict_left", "restrictright": "restrict_right"}
_IGNORED = {"forest", "array", "array2", "array1", "arrayh", "hybrid", "arrayhf", "hybridfast", "ways", "ordered", "ord", "showspeed",
            "ss", "prealloc", "preallocate", "initialsize", "nzo", "nonzeroonly", "statscolumns", "columns", "cols", "threads", "t"}
_FILE_ALIASES = {"in": "in1", "in1": "in1", "in2": "in2", "ref": "ref", "literal": "literal",
                 "out": "outm1", "out1": "outm1", "outm": "outm1", "outm1": "outm1", "outmatched": "outm1", "outmatched1": "outm1",
                 "out2": "outm2", "outm2": "outm2", "outmatched2": "outm2",
                 "outu": "outu1", "outu1": "outu1", "outunmatched": "outu1", "outunmatched1": "outu1",
                 "outu2": "outu2", "outunmatched2": "outu2", "stats": "stats", "scafstats": "stats", "refstats": "refstats"}  # jgi/Seal.java:169-190
_FILES = set(_FILE_ALIASES)


def parse_seal_args(args, device=0):
    """seal.sh key=value flags -> (SealCfg, {file flags}). Same names, aliases and defaults as jgi/Seal.java:150-380;
    flags of paths the device does not serve raise ValueError instead of being ignored."""
    kw = {}
    files = {}
    for arg in args:
        a, _, b = arg.partition("=")
        a = a.lower()
        b = b if _ else None
        if b is not None and b.lower() == "null":
            b = None
        if a in _INT_FLAGS:
            kw[_INT_FLAGS[a]] = int(b)
        elif a in ("minkmerfraction", "minfraction", "mkf"):
            kw["min_kmer_fraction"] = float(b)
        elif a in ("mm", "maskmiddle"):
            if b is None or b[0].isalpha():  # :255-261
                kw["mask_middle"] = 1 if _parse_bool(b) else 0
            else:
                kw["mid_mask_len"] = int(b)
                kw["mask_middle"] = 1 if int(b) > 0 else 0
        elif a == "rcomp":
            kw["rcomp"] = 1 if _parse_bool(b) else 0
        elif a in ("forbidns", "forbidn", "fn"):
            kw["forbid_ns"] = 1 if _parse_bool(b) else 0
        elif a in ("ambiguous", "ambig"):
            if b is None or b.lower() not in _AMBIG:
                raise ValueError(arg)
            kw["ambig_mode"] = _AMBIG[b.lower()]
        elif a in ("match", "mode"):
            if b is None or b.lower() not in _MATCH:
                raise ValueError(arg)
            kw["match_mode"] = _MATCH[b.lower()]
        elif a in ("findbestmatch", "fbm"):
            kw["match_mode"] = MATCH_ALL if _parse_bool(b) else MATCH_FIRST
        elif a in ("firstuniquematch", "fum"):
            if _parse_bool(b):
                kw["match_mode"] = MATCH_UNIQUE
        elif a in ("keeppairstogether", "kpt"):
            kw["keep_pairs_together"] = 1 if _parse_bool(b) else 0
        elif a in ("clearzone", "cz"):
            if "." in b:  # :357-362
                kw["clearzone_fraction"] = float(b)
            else:
                kw["clearzone"] = int(b)
        elif a in ("clearzonefraction", "czf"):
            kw["clearzone_fraction"] = float(b)
        elif a in ("qhdist", "queryhammingdistance", "edits", "edist", "editdistance"):
            if int(b) != 0:
                raise ValueError(f"{a}={b}: not served by the device path (include/seal_b200.h)")
        elif a in ("processcontainedref", "countvector", "trackbarcodes", "ecco", "ecc", "rename"):
            if _parse_bool(b):
                raise ValueError(f"{a}: not served by the device path (include/seal_b200.h)")
        elif a in _FILES:
            files[_FILE_ALIASES[a]] = b
        elif a in _IGNORED:
            pass
        else:
            raise ValueError(f"unknown flag: {arg}")
    k = kw.get("k", 31)
    if not 0 < k < 32:
        raise ValueError("k must be at least 1; default is 31.")  # :225
    if not 0 <= kw.get("hdist", 0) < 4:
        raise ValueError("hamming distance must be between 0 and 3; default is 0.")  # :229
    if not 0 <= kw.get("speed", 0) <= 16:
        raise ValueError("Speed range is 0 to 16.")  # :243
    return make_cfg(device=device, **kw), files


def process_sharded(engine, bases, offsets, paired, first_numeric_id=0, group=None):
    """N ranks, one engine (SealIndexGPU built from the same reference) per rank: this rank matches its contiguous slice
    of the batch, the per-unit results come back in input order on every rank, totals and per-reference counters are
    summed -- as Seal sums its ProcessThreads' counters (jgi/Seal.java:1640-1680). No per-read collective.
    -> (fields dict, stats dict, [reads, bases, frags, ambig] of THIS call)."""
    import torch
    import torch.distributed as dist

    from .shard import gather_in_order, shard_reads, sum_stats
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bases = np.ascontiguousarray(bases, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    r0, r1, loff = shard_reads(offsets, paired, world, rank)
    before = [x.copy() for x in engine.scaffold_counts()]
    first = first_numeric_id + (r0 // 2 if paired else r0)  # Read.numericID counts pairs
    res, st = engine.process(bases[offsets[r0]:offsets[r1]], loff, paired, first)
    stride = max(res.stride, 0)
    n_local = len(res.n_assigned)
    local = {k: v for k, v in res.fields().items() if k != "ids"}
    local["ids"] = res.ids[:n_local * stride]
    merged = gather_in_order(local, group)
    total = sum_stats(st.as_dict(), group)
    delta = np.stack([a - b for a, b in zip(engine.scaffold_counts(), before)])
    t = torch.from_numpy(delta)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    return merged, total, list(t.cpu().numpy())


def format_stats(names, stats, counts, in1, in2=None, columns=5, nonzero_only=True):
    """The `stats=` file of seal.sh (jgi/Seal.java:899-950 writeStats): one line per reference sequence, sorted by
    bases, then reads (both descending), then name (structures/StringCount.java:36-40); percentages with five decimals.
    names[i] belongs to id i+1; stats = seal_stats as a dict; counts = [reads, bases, frags, ambig] by id."""
    reads, bases, _, ambig = counts
    rows = []
    asum = 0
    for i, name in enumerate(names, start=1):
        if reads[i] > 0 or not nonzero_only:
            asum += int(ambig[i])
            rows.append((name, int(reads[i]), int(bases[i]), int(ambig[i])))
    rows.sort(key=lambda r: r[0])
    rows.sort(key=lambda r: (-r[2], -r[1]))  # stable: ties keep the name order
    rmult = 100.0 / (stats["reads_in"] if stats["reads_in"] > 0 else 1)
    bmult = 100.0 / (stats["bases_in"] if stats["bases_in"] > 0 else 1)
    out = ["#File\t" + in1 + ("" if in2 is None else "\t" + in2) + "\n"]
    if columns == 3:
        out.append("#Total\t%d\n" % stats["reads_in"])
        out.append("#Matched\t%d\t%.5f%%\n" % (stats["reads_matched"], rmult * stats["reads_matched"]))
        out.append("#Name\tReads\tReadsPct\n")
        out += ["%s\t%d\t%.5f%%\n" % (n, r, r * rmult) for n, r, _, _ in rows]
    else:
        out.append("#Total\t%d\t%d\n" % (stats["reads_in"], stats["bases_in"]))
        # the reference's format string consumes three of its five arguments (:941)
        out.append("#Matched\t%d\t%.5f%%\t%d\n" % (stats["reads_matched"], rmult * stats["reads_matched"], stats["bases_matched"]))
        out.append("#Name\tReads\tReadsPct\tBases\tBasesPct\tAmbigReads\n")
        out += ["%s\t%d\t%.5f%%\t%d\t%.5f%%\t%d\n" % (n, r, r * rmult, b, b * bmult, a) for n, r, b, a in rows]
    return "".join(out)


class Seal:
    """seal.sh from files: `Seal(["in=r1.fq", "in2=r2.fq", "ref=a.fa,b.fa", "outm=m.fq", "outu=u.fq", "stats=s.txt", ...])`.
    FASTQ in / FASTQ out through the native feed (include/fastq_b200.h), streamed in bounded blocks: whole records, the same
    number from both mate files; a pair goes to outm when it was assigned to at least one reference, else to outu
    (jgi/Seal.java:2278-2286); Read.numericID runs over the whole input, as ambig=random needs it (:2403).
    `engine` (tests): anything with SealIndexGPU's add_ref / finalize / process / scaffold_counts."""

    def __init__(self, args, device=0, engine=None):
        from .fasta import read_fasta
        self.cfg, self.files = parse_seal_args(args, device)
        if not self.files.get("in1"):
            raise ValueError("in= is required")
        self.engine = engine(self.cfg) if engine is not None else SealIndexGPU(self.cfg)
        self.names = []
        self.ref_files, self.scaf_per_file, self.scaf_lengths = [], [], []
        for path in (self.files.get("ref") or "").split(","):
            if path:
                names, bases, offsets = read_fasta(path)
                self.names += names
                self.ref_files.append(path)
                self.scaf_per_file.append(len(names))
                self.scaf_lengths += [int(x) for x in np.diff(offsets)]
                self.engine.add_ref(bases, offsets)
        for i, lit in enumerate((self.files.get("literal") or "").split(",")):
            if lit:
                self.names.append(f"literal_{i}")
                b = np.frombuffer(lit.encode(), np.uint8)
                self.ref_files.append("literal")
                self.scaf_per_file.append(1)
                self.scaf_lengths.append(len(b))
                self.engine.add_ref(b, np.array([0, len(b)], np.int64))
        self.stored, self.entries, self.ref_kmers = self.engine.finalize()
        self.stats = None

    def process(self, block_bytes=None):
        import os

        from .fastq import FastqBatch, iter_fastq_blocks
        f = self.files
        paired = bool(f.get("in2"))
        per = 2 if paired else 1
        block = int(block_bytes or os.environ.get("BBDUK_B200_FEED_BLOCK", 256 << 20))
        routes = [(matched, path, sel) for matched, p1, p2 in ((True, f.get("outm1"), f.get("outm2")), (False, f.get("outu1"), f.get("outu2")))
                  if p1 for path, sel in (((p1, 1), (p2, 2)) if p2 else ((p1, 0),))]
        sinks = {path: open(path, "wb") for _, path, _ in routes}
        total = {n: 0 for n, _ in SealStats._fields_ if n != "reserved"}
        first_id = 0
        try:
            for text1, text2 in iter_fastq_blocks(f["in1"], f.get("in2") or None, block, unit=per):
                fb = FastqBatch(text1, text2)
                bases, offsets = fb.arrays()
                n = len(offsets) - 1
                res, st = self.engine.process(bases, offsets, paired, first_id)
                first_id += n // per
                for k, v in st.as_dict().items():
                    total[k] += v
                if paired and not self.cfg.keep_pairs_together:  # assigned = both mates' sites (:2270-2272)
                    hit = (res.n_assigned[0::2] + res.n_assigned[1::2]) > 0
                else:
                    hit = res.n_assigned > 0
                flags = np.repeat(np.where(hit, 2, 0).astype(np.uint8), per)  # BBDUK_F_REMOVED marks the matched units
                lens = np.diff(offsets).astype(np.int32)
                for matched, path, sel in routes:
                    fb.format(per, np.zeros(n, np.int32), lens, flags, removed=matched, mate_sel=sel, trim_removed=True).tofile(sinks[path])
        finally:
            for fh in sinks.values():
                fh.close()
        self.stats = total
        self.counts = self.engine.scaffold_counts()
        if f.get("stats"):
            with open(f["stats"], "w") as fh:
                fh.write(format_stats(self.names, total, self.counts, f["in1"], f.get("in2")))
        if f.get("refstats"):
            with open(f["refstats"], "w") as fh:
                fh.write(format_refstats(self.ref_files, self.scaf_per_file, np.array(self.scaf_lengths, np.int64), total, self.counts,
                                         f["in1"], f.get("in2")))
        return total


_EXTENSIONS = ("fa", "fasta", "fna", "ffn", "frn", "fsa", "fas", "seq", "faa", "fq", "fastq", "txt", "gz", "bz2", "zip", "xz", "zst")


def strip_to_core(path):
    """fileIO/ReadWrite.java:1896-1922 stripToCore for the extensions a reference file carries (a subset of FileFormat.EXTENSION_LIST)."""
    name = path.replace("\\", "/").rsplit("/", 1)[-1]
    while True:
        for ext in _EXTENSIONS:
            if name.endswith("." + ext):
                name = name[:-len(ext) - 1]
                break
        else:
            return name


def format_refstats(ref_files, scaf_per_file, scaf_lengths, stats, counts, in1, in2=None, nonzero_only=True):
    """The `refstats=` file (jgi/Seal.java:1031-1096 writeRefStats): the per-sequence counters summed per reference FILE,
    coverage, RPKM and FPKM (single-precision multiplier 1e9f / mapped, double-precision 1 / length).
    scaf_per_file[i] sequences of ref_files[i] in id order; scaf_lengths[id-1] their lengths."""
    reads, bases, frags, ambig = counts
    mapped = int(np.sum(reads))
    mult = np.float32(1000000000.0) / np.float32(max(1, mapped))
    out = ["#File\t" + in1 + ("" if in2 is None else "\t" + in2) + "\n", "#Reads\t%d\n" % stats["reads_in"], "#Mapped\t%d\n" % mapped,
           "#References\t%d\n" % len(ref_files), "#Name\tLength\tScaffolds\tBases\tCoverage\tReads\tRPKM\tFrags\tFPKM\tAmbigReads\n"]
    sid = 1
    for path, n in zip(ref_files, scaf_per_file):
        r = int(np.sum(reads[sid:sid + n]))
        b = int(np.sum(bases[sid:sid + n]))
        f = int(np.sum(frags[sid:sid + n]))
        a = int(np.sum(ambig[sid:sid + n]))
        ln = int(np.sum(scaf_lengths[sid - 1:sid - 1 + n]))
        sid += n
        invlen = 1.0 / max(1, ln)
        mult2 = float(mult) * invlen
        if r > 0 or not nonzero_only:
            out.append("%s\t%d\t%d\t%d\t%.4f\t%d\t%.4f\t%d\t%.4f\t%d\n" % (strip_to_core(path), ln, n, b, b * invlen, r, r * mult2, f, f * mult2, a))
    return "".join(out)
