"""Host-side mirror of the reference's BBDuk surface for the k-mer match-and-trim path.

`parse_args` accepts the bbduk.sh `key=value` flags that reach the path (same names, aliases and
defaults as jgi/BBDuk.java:186-560 and parse/Parser.java:487-506) and produces the POD config of the
C ABI. `BBDukIndexGPU` is the fourth index implementation SURVEY.md 8b describes (next to
bbduk/BBDukIndexMod|Mask|Mask2): it owns a device table and answers whole batches.
`BBDuk` wires them into the reads-in / reads-out tool for FASTQ files.

Everything numeric happens in libbbduk_b200.so; this file is plumbing (no CPU fallback exists).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._abi import (BBDukOut, BBDukStats, BBDukTableDesc, F_DISCARDED, F_REMOVED, F_SPLIT, GEN_JGI, GEN_S, Outputs,
                   default_cfg)
from .fasta import pack, read_fasta, read_fastq


def _parse_boolean(s):
    """parse/Parse.java:187-195"""
    if s is None or len(s) < 1:
        return True
    if len(s) == 1:
        return s.lower() in ("t", "1")
    if s.lower() in ("null", "none"):
        return False
    return s.lower() == "true"


def _parse_poly(b):
    """parse/Parser.java:426-437"""
    if b is None:
        return 2
    if b[:1].isdigit():
        return int(b)
    return 2 if _parse_boolean(b) else 0


def parse_args(args, generation=GEN_JGI):
    """bbduk.sh-style argv -> (bbduk_cfg, io dict). Unknown keys raise, like the reference
    ("Unknown parameter", jgi/BBDuk.java:561)."""
    cfg = default_cfg()
    cfg.generation = generation
    io = {"in1": None, "in2": None, "out1": None, "out2": None, "outm1": None, "outm2": None, "ref": [],
          "literal": [], "stats": None, "interleaved": None, "ordered": generation == GEN_S, "ottm": False,
          "tbo": False, "strictoverlap": True, "minoverlap": -1, "mininsert": -1,
          "qtrim_left": False, "qtrim_right": False, "trimq": 6.0, "mbq": 0, "maxns": -1, "maxlen": 0,
          "trimpolya": 0, "trimpolygleft": 0, "trimpolygright": 0, "filterpolyg": 0, "trimpolycleft": 0, "trimpolycright": 0,
          "filterpolyc": 0, "maxnonpoly": 1, "entropy": -1.0, "entropyk": 5, "entropywindow": 50, "entropymask": 0, "entropytrim": False,
          "maq": 0.0, "maqb": 0, "maxnrate": 1.0, "mcb": 0, "minbasefrequency": 0.0,
          "trim_mode": 0, "window_length": 4, "min_good_interval": 2}
    for arg in args:
        sp = arg.split("=")
        a = sp[0].lower()
        b = sp[1] if len(sp) > 1 else None
        if a in ("in", "in1"):
            io["in1"] = b
        elif a == "in2":
            io["in2"] = b
        elif a in ("out", "out1", "outu", "outu1", "outnonmatch", "outnonmatch1"):
            io["out1"] = b
        elif a in ("out2", "outu2", "outnonmatch2"):
            io["out2"] = b
        elif a in ("outb", "outm", "outb1", "outm1", "outbad", "outbad1", "outmatch", "outmatch1"):
            io["outm1"] = b
        elif a in ("outb2", "outm2", "outbad2", "outmatch2"):
            io["outm2"] = b
        elif a in ("stats", "scafstats"):
            io["stats"] = b
        elif a in ("ref", "adapters"):
            io["ref"] = [] if b is None else b.split(",")
        elif a == "literal":
            io["literal"] = [] if b is None else b.split(",")
        elif a in ("interleaved", "int"):
            io["interleaved"] = _parse_boolean(b)
        elif a in ("ordered", "ord"):
            io["ordered"] = _parse_boolean(b)
        elif a in ("ottm", "outputtrimmedtomatch"):
            io["ottm"] = _parse_boolean(b)
        elif a in ("t", "threads", "overwrite", "ow", "showspeed", "ss", "prealloc", "preallocate"):
            pass  # host runtime knobs with no effect on results
        elif a == "skipr1":
            cfg.skip_r1 = _parse_boolean(b)
        elif a == "skipr2":
            cfg.skip_r2 = _parse_boolean(b)
        elif a == "k":
            cfg.k = int(b)
        elif a in ("mink", "kmin"):
            cfg.mink = int(b)
        elif a in ("useshortkmers", "shortkmers", "usk"):
            cfg.use_short_kmers = _parse_boolean(b)
        elif a in ("trimextra", "trimpad", "tp"):
            cfg.trim_pad = int(b)
        elif a in ("hdist", "hammingdistance"):
            cfg.hdist = int(b)
        elif a in ("qhdist", "queryhammingdistance"):
            cfg.qhdist = int(b)
        elif a in ("edits", "edist", "editdistance"):
            cfg.edist = int(b)
        elif a in ("hdist2", "hammingdistance2"):
            cfg.hdist2 = int(b)
        elif a in ("qhdist2", "queryhammingdistance2"):
            cfg.qhdist2 = int(b)
        elif a in ("edits2", "edist2", "editdistance2"):
            cfg.edist2 = int(b)
        elif a in ("maxskip", "maxrskip", "mxs"):
            cfg.max_skip = int(b)
        elif a in ("minskip", "minrskip", "mns"):
            cfg.min_skip = int(b)
        elif a in ("skip", "refskip", "rskip"):
            cfg.min_skip = cfg.max_skip = int(b)
        elif a == "qskip":
            cfg.qskip = int(b)
        elif a == "speed":
            cfg.speed = int(b)
        elif a in ("maxbadkmers", "mbk"):
            cfg.max_bad_kmers = int(b)
        elif a in ("minhits", "minkmerhits", "mkh"):
            cfg.max_bad_kmers = int(b) - 1
        elif a in ("minkmerfraction", "minfraction", "mkf"):
            cfg.min_kmer_fraction = float(b)
        elif a in ("mincoveredfraction", "mincovfraction", "mcf"):
            cfg.min_covered_fraction = float(b)
        elif a in ("mm", "maskmiddle"):
            if b is None or b[:1].isalpha():
                cfg.mask_middle = _parse_boolean(b)
                if not cfg.mask_middle:  # `mm=5 mm=f` ends with maskMiddle=false and midMaskLen=0 (bbduk/BBDukParser.java:232-236)
                    cfg.mid_mask_len = 0
            else:
                cfg.mid_mask_len = int(b)
                cfg.mask_middle = int(b) > 0
        elif a == "rcomp":
            cfg.rcomp = _parse_boolean(b)
        elif a in ("forbidns", "forbidn", "fn"):
            cfg.forbid_ns = _parse_boolean(b)
        elif a in ("findbestmatch", "fbm"):
            cfg.find_best_match = _parse_boolean(b)
        elif a == "kfilter":
            if _parse_boolean(b):
                cfg.ktrim_left = cfg.ktrim_right = cfg.ktrim_n = cfg.ksplit = 0
        elif a == "ksplit":
            if _parse_boolean(b):
                cfg.ksplit = 1
                cfg.ktrim_left = cfg.ktrim_right = cfg.ktrim_n = 0
            else:
                cfg.ksplit = 0
        elif a == "ktrim":
            v = (b or "").lower()
            if v in ("rl", "lr", "tips"):
                cfg.ktrim_left = cfg.ktrim_right = 1
                cfg.ktrim_n = cfg.ksplit = 0
            elif v in ("left", "l"):
                cfg.ktrim_left, cfg.ktrim_right, cfg.ktrim_n, cfg.ksplit = 1, 0, 0, 0
            elif v in ("right", "r"):
                cfg.ktrim_left, cfg.ktrim_right, cfg.ktrim_n, cfg.ksplit = 0, 1, 0, 0
            elif v == "n":
                cfg.ktrim_left, cfg.ktrim_right, cfg.ktrim_n, cfg.ksplit = 0, 0, 1, 0
            elif len(v) == 1 and v not in ("t", "f"):
                cfg.ktrim_left, cfg.ktrim_right, cfg.ktrim_n, cfg.ksplit = 0, 0, 1, 0
                cfg.trim_symbol = ord(b[0])
            else:
                if v not in ("f", "false"):
                    raise ValueError("Invalid setting for ktrim - values must be f (false), l (left), r (right), "
                                     "rl (tips), or n.")
                cfg.ktrim_left = cfg.ktrim_right = 0
        elif a in ("trimtips", "ktrimtips"):
            if b:
                cfg.ktrim_left = cfg.ktrim_right = 1
                cfg.ktrim_n = cfg.ksplit = 0
                cfg.restrict_left = cfg.restrict_right = int(b)
        elif a in ("kmask", "mask"):
            if b is not None and b.lower() in ("lc", "lowercase"):
                cfg.kmask_lowercase = 1
                cfg.ktrim_left, cfg.ktrim_right, cfg.ktrim_n, cfg.ksplit = 0, 0, 1, 0
            else:
                if _parse_boolean(b):
                    b = "N"
                if b is not None and len(b) == 1:
                    cfg.ktrim_left, cfg.ktrim_right, cfg.ktrim_n = 0, 0, 1
                    cfg.trim_symbol = ord(b)
                else:
                    cfg.ktrim_n = _parse_boolean(b)
        elif a in ("kmaskfullycovered", "maskfullycovered", "mfc"):
            cfg.kmask_fully_covered = _parse_boolean(b)
        elif a == "ktrimright":
            cfg.ktrim_right = _parse_boolean(b)
            cfg.ktrim_left = cfg.ktrim_n = not cfg.ktrim_right
        elif a == "ktrimleft":
            cfg.ktrim_left = _parse_boolean(b)
            cfg.ktrim_right = cfg.ktrim_n = not cfg.ktrim_left
        elif a == "ktrimn":
            cfg.ktrim_n = _parse_boolean(b)
            cfg.ktrim_left = cfg.ktrim_right = not cfg.ktrim_n
        elif a == "ktrimexclusive":
            cfg.ktrim_exclusive = _parse_boolean(b)
        elif a in ("tpe", "tbe", "trimpairsevenly"):
            cfg.trim_pairs_evenly = _parse_boolean(b)
        elif a == "restrictleft":
            cfg.restrict_left = int(b)
        elif a == "restrictright":
            cfg.restrict_right = int(b)
        elif a in ("ml", "minlen", "minlength"):
            cfg.min_read_length = int(b)
        elif a in ("mlf", "minlenfrac", "minlenfraction", "minlengthfraction"):
            cfg.min_len_fraction = float(b)
        elif a in ("requirebothbad", "rbb"):
            cfg.require_both_bad = _parse_boolean(b)
        elif a in ("removeifeitherbad", "rieb"):
            cfg.require_both_bad = not _parse_boolean(b)
        elif a in ("trimfailures", "trimfailuresto1bp"):
            cfg.trim_failures_to_1bp = _parse_boolean(b)
        elif a in ("tbo", "trimbyoverlap"):  # jgi/BBDuk.java:380-393
            io["tbo"] = _parse_boolean(b)
        elif a == "strictoverlap":
            io["strictoverlap"] = _parse_boolean(b)
        elif a == "minoverlap":
            io["minoverlap"] = int(b)
        elif a == "mininsert":
            io["mininsert"] = int(b)
        elif a == "qtrim":  # parse/Parser.java:347-367
            v = (b or "").lower()
            if v == "":
                io["qtrim_left"] = io["qtrim_right"] = True
            elif v in ("left", "l"):
                io["qtrim_left"], io["qtrim_right"] = True, False
            elif v in ("right", "r"):
                io["qtrim_left"], io["qtrim_right"] = False, True
            elif v in ("both", "rl", "lr"):
                io["qtrim_left"] = io["qtrim_right"] = True
            elif v in ("window", "w") or v.startswith("window,") or v.startswith("w,"):  # right end only, TrimRead.windowMode
                io["qtrim_left"], io["qtrim_right"] = False, True
                io["trim_mode"] = 1
                parts = v.split(",")
                if len(parts) > 1:
                    io["window_length"] = int(parts[1])
            elif v[:1].isdigit():  # qtrim=<number> sets trimq and trims the right end (parse/Parser.java:364-366)
                io["trimq"] = float(v)
                io["qtrim_right"] = True
            else:
                io["qtrim_left"] = io["qtrim_right"] = _parse_boolean(b)
        elif a in ("optitrim", "otf", "otm"):  # parse/Parser.java:368-375 (a numeric optimalBias is not supported here)
            if b and (b[0] == "." or b[0].isdigit()):
                raise NotImplementedError("optitrim=<bias> (TrimRead.optimalBias) is not on the device path")
            if _parse_boolean(b):
                io["trim_mode"] = 0
            elif io["trim_mode"] == 0:
                io["trim_mode"] = 2
        elif a == "trimgoodinterval":
            io["min_good_interval"] = int(b)
        elif a in ("trimright", "qtrimright"):
            io["qtrim_right"] = _parse_boolean(b)
        elif a in ("trimleft", "qtrimleft"):
            io["qtrim_left"] = _parse_boolean(b)
        elif a in ("trimq", "trimquality"):
            io["trimq"] = float(b)
        elif a in ("minavgquality", "minaveragequality", "maq"):  # parse/Parser.java:516-526
            if "," in b:
                q_, n_ = b.split(",")
                io["maq"], io["maqb"] = float(q_), int(n_)
            else:
                io["maq"] = float(b)
        elif a in ("minavgqualitybases", "maqb"):
            io["maqb"] = int(b)
        elif a in ("minbasequality", "mbq"):
            io["mbq"] = int(b)
        elif a == "maxns":
            io["maxns"] = int(b)
        elif a in ("maxnrate", "maxnfraction"):  # parse/Parser.java:510-513: a fraction, or a percentage when above 1
            v = float(b)
            io["maxnrate"] = v / 100 if v > 1 else (v if v >= 0 else 1.0)  # unset / negative = 1 = off (jgi/BBDuk.java:629)
        elif a in ("minconsecutivebases", "mcb"):
            io["mcb"] = int(b)
        elif a == "minbasefrequency":  # jgi/BBDuk.java:465-466
            io["minbasefrequency"] = float(b)
        elif a in ("maxlength", "maxreadlength", "maxreadlen", "maxlen"):
            io["maxlen"] = int(b)
        elif a in ("trimpolya", "trimpolygleft", "trimpolygright", "filterpolyg", "trimpolycleft", "trimpolycright", "filterpolyc",
                   "maxnonpoly"):  # parse/Parser.java:386-437
            io[a] = _parse_poly(b)
        elif a == "trimpolyg":
            io["trimpolygleft"] = io["trimpolygright"] = _parse_poly(b)
        elif a == "trimpolyc":
            io["trimpolycleft"] = io["trimpolycright"] = _parse_poly(b)
        elif a in ("minentropy", "entropy", "entropyfilter"):  # jgi/BBDuk.java:418-419
            io["entropy"] = float(b)
        elif a in ("entropymask", "maskentropy"):  # jgi/BBDuk.java:422-432: 1 = to N, 2 = to lower case
            v = (b or "").lower()
            io["entropymask"] = 1 if b is None else 2 if v in ("lc", "lowercase") else 0 if v == "filter" else int(_parse_boolean(b))
        elif a in ("entropytrim", "trimentropy"):  # jgi/BBDuk.java:433-434, parseEnd :4846-4852 (any end = both ends, :4448-4478)
            v = (b or "rl").lower()
            io["entropytrim"] = v in ("right", "r", "left", "l", "rl", "lr", "both", "b") or _parse_boolean(b)
        elif a in ("entropyk", "ek"):  # parse/Parser.java:955-960
            io["entropyk"] = int(b)
        elif a in ("entropywindow", "ew"):
            io["entropywindow"] = int(b)
        elif a in ("entropymask", "maskentropy", "entropytrim", "trimentropy", "entropymark", "markentropy"):
            if b is None or _parse_boolean(b) or b[:1].isalpha() and b.lower() not in ("f", "false"):
                raise NotImplementedError(f"{a} is not on the device path (only the entropy= read filter is)")
        elif a == "usequality":
            if _parse_boolean(b):
                raise NotImplementedError("usequality=t (quality-weighted overlap) is not on the device path")
        else:
            raise ValueError(f"Unknown parameter {arg}")
    return cfg, io


class BBDukIndexGPU:
    """Device-resident k-mer index + batched per-read k-mer block (the C ABI behind a Python handle)."""

    def __init__(self, cfg, _handle=None):
        self.lib = _lib.load()
        self.cfg = cfg
        if _handle is not None:  # a handle the library created itself (bbduk_b200_replicate)
            self.h = _handle
        else:
            h = C.c_void_p()
            rc = self.lib.bbduk_b200_create(C.byref(cfg), C.byref(h))
            if rc:
                raise RuntimeError("bbduk_b200_create: " + self.lib.bbduk_b200_last_error(None).decode())
            self.h = h
        self.stored_kmers = None
        self.n_scaffolds = 0

    TRANSPORT = {0: "built", 1: "nccl", 2: "peer"}

    def replicate(self, device_ids):
        """In-library replication of the finished table to the listed GPUs of THIS process (bbduk_b200_replicate:
        one NCCL broadcast per blob, or peer copies) -> one BBDukIndexGPU per device id."""
        ids = (C.c_int32 * len(device_ids))(*device_ids)
        hs = (C.c_void_p * len(device_ids))()
        self._check(self.lib.bbduk_b200_replicate(self.h, ids, len(device_ids), hs), "replicate")
        out = []
        for i, dev in enumerate(device_ids):
            e = BBDukIndexGPU(self.cfg, _handle=C.c_void_p(hs[i]))
            e.stored_kmers, e.n_scaffolds, e.device = self.stored_kmers, self.n_scaffolds, dev
            out.append(e)
        return out

    @property
    def transport(self):
        return self.TRANSPORT.get(int(self.lib.bbduk_b200_replica_transport(self.h)), "?")

    @staticmethod
    def process_sharded(engines, bases, offsets, paired, want_mask=False, out=None):
        """One batch cut into len(engines) contiguous slices, slice i on engines[i] (its own GPU and host thread);
        results in input order, counters summed (bbduk_b200_process_sharded)."""
        lib = engines[0].lib
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        if out is None:
            out = Outputs(n, np.diff(offsets), want_mask=want_mask)
        st = BBDukStats()
        o = out.struct()
        hs = (C.c_void_p * len(engines))(*[e.h for e in engines])
        engines[0]._check(lib.bbduk_b200_process_sharded(hs, len(engines), bases.ctypes.data, offsets.ctypes.data, n,
                                                         int(bool(paired)), C.byref(o), C.byref(st)), "process_sharded")
        return out, st

    @staticmethod
    def scaffold_counts_sum(engines):
        n = engines[0].n_scaffolds + 1
        rc_, bc = np.zeros(n, np.int64), np.zeros(n, np.int64)
        hs = (C.c_void_p * len(engines))(*[e.h for e in engines])
        engines[0]._check(engines[0].lib.bbduk_b200_scaffold_counts_sum(hs, len(engines), rc_.ctypes.data, bc.ctypes.data, n),
                          "scaffold_counts_sum")
        return rc_, bc

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: " + self.lib.bbduk_b200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.bbduk_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- table ---------------------------------------------------------------------------------
    def add_ref(self, bases, offsets):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        self._check(self.lib.bbduk_b200_add_ref(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1),
                    "add_ref")
        self.n_scaffolds += len(offsets) - 1

    def finalize(self):
        n = C.c_int64(0)
        self._check(self.lib.bbduk_b200_finalize(self.h, C.byref(n)), "finalize")
        self.stored_kmers = int(n.value)
        return self.stored_kmers

    def table_describe(self):
        d = BBDukTableDesc()
        self._check(self.lib.bbduk_b200_table_describe(self.h, C.byref(d)), "table_describe")
        return d

    @staticmethod
    def _view(ptr, nbytes):
        """zero-copy uint8 torch view of library-owned device memory (CUDA array interface)"""
        import torch

        class _Mem:
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                      "version": 2}
        return torch.as_tensor(m, device=torch.device("cuda", torch.cuda.current_device()))

    def dump_table(self):
        """(sorted keys, ids) copied back from the device, for table-parity tests."""
        d = self.table_describe()
        k = self._view(d.d_keys, d.n_slots * 8).cpu().numpy().view(np.uint64)
        v = self._view(d.d_vals, d.n_slots * 4).cpu().numpy().view(np.int32)
        m = k != np.uint64(0xFFFFFFFFFFFFFFFF)
        k, v = k[m], v[m]
        order = np.argsort(k)
        return k[order], v[order]

    def broadcast_table(self, src=0, group=None):
        """Replicate rank `src`'s finished table to every rank: one NCCL broadcast per blob (keys, ids,
        filter image) straight out of / into the library's device arrays over NVLink.
        Non-source ranks must NOT have called finalize."""
        import torch
        import torch.distributed as dist
        rank = dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        geo = torch.zeros(12, dtype=torch.int64, device=dev)
        if rank == src:
            d = self.table_describe()
            geo[:4] = torch.tensor([d.n_slots, d.n_filter_words, d.stored_kmers, d.n_scaffolds])
            geo[4:12] = torch.tensor(list(d.scalars))
        dist.broadcast(geo, src, group=group)
        g = geo.cpu().tolist()
        if rank != src:
            d = BBDukTableDesc()
            d.n_slots, d.n_filter_words, d.stored_kmers, d.n_scaffolds = g[0], g[1], g[2], int(g[3])
            for i in range(8):
                d.scalars[i] = g[4 + i]
            self._check(self.lib.bbduk_b200_table_alloc(self.h, C.byref(d)), "table_alloc")
        for ptr, nbytes in ((d.d_keys, g[0] * 8), (d.d_vals, g[0] * 4), (d.d_filter, g[1] * 4)):
            dist.broadcast(self._view(ptr, nbytes), src, group=group)
        torch.cuda.synchronize()
        if rank != src:
            self._check(self.lib.bbduk_b200_table_commit(self.h), "table_commit")
            self.stored_kmers = g[2]
            self.n_scaffolds = int(g[3])
        return g[2]

    # -- reads ---------------------------------------------------------------------------------
    def process(self, bases, offsets, paired, want_mask=False, out=None):
        """HOST buffers in, host struct-of-arrays out (bbduk_b200_process)."""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        if out is None:
            out = Outputs(n, np.diff(offsets), want_mask=want_mask)
        st = BBDukStats()
        o = out.struct()
        self._check(self.lib.bbduk_b200_process(self.h, bases.ctypes.data, offsets.ctypes.data, n, int(bool(paired)),
                                                C.byref(o), C.byref(st)), "process")
        return out, st

    def pack(self, bases):
        """the 2-bit stream F and the defined bits D of concatenated ASCII bases (bbduk_b200_pack_bases), as a packing parser would hold them"""
        bases = np.ascontiguousarray(bases, np.uint8)
        g = (len(bases) + 15) // 16
        F = np.zeros(g + 16, np.uint32)
        D = np.zeros(g + 32, np.uint16)
        self._check(self.lib.bbduk_b200_pack_bases(bases.ctypes.data, len(bases), F.ctypes.data, D.ctypes.data), "pack_bases")
        return F, D

    def process_packed(self, F, D, offsets, paired, out=None):
        """HOST 2-bit stream + defined bits in, host struct-of-arrays out (bbduk_b200_process_packed)."""
        F = np.ascontiguousarray(F, np.uint32)
        D = np.ascontiguousarray(D, np.uint16)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        if out is None:
            out = Outputs(n, np.diff(offsets))
        st = BBDukStats()
        o = out.struct()
        self._check(self.lib.bbduk_b200_process_packed(self.h, F.ctypes.data, D.ctypes.data, offsets.ctypes.data, n,
                                                       int(bool(paired)), C.byref(o), C.byref(st)), "process_packed")
        return out, st

    def process_device(self, d_bases, d_offsets, n_reads, paired, d_out, d_stats=None, stream=None):
        """DEVICE buffers (torch tensors or raw pointers); d_out: dict name -> tensor/pointer."""
        def ptr(x):
            if x is None:
                return None
            return x.data_ptr() if hasattr(x, "data_ptr") else int(x)
        o = BBDukOut()
        for name in ("id0", "id0b", "lo", "hi", "flags", "count", "maskbits", "mask_off"):
            setattr(o, name, ptr(d_out.get(name)))
        self._check(self.lib.bbduk_b200_process_device(self.h, ptr(d_bases), ptr(d_offsets), n_reads,
                                                       int(bool(paired)), C.byref(o), ptr(d_stats), ptr(stream)),
                    "process_device")

    # -- trim by overlap (tbo=t), the step after the k-mer block (jgi/BBDuk.java:2878-2926) -----------
    def tbo_cfg(self, **kw):
        from ._abi import BBDukTboCfg
        c = BBDukTboCfg()
        self.lib.bbduk_b200_tbo_cfg_default(C.byref(c))
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def tbo(self, bases, quals, offsets, out, cfg=None):
        """HOST buffers; `out` is the Outputs of process(): hi / flags are updated in place.
        -> (insert per pair, [reads trimmed, bases trimmed])"""
        cfg = cfg or self.tbo_cfg()
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        q = None if quals is None else np.ascontiguousarray(quals, np.uint8)
        n = len(offsets) - 1
        insert = np.full(n // 2, -9, np.int32)
        st = np.zeros(2, np.int64)
        self._check(self.lib.bbduk_b200_tbo(self.h, C.byref(cfg), bases.ctypes.data, None if q is None else q.ctypes.data,
                                            offsets.ctypes.data, n, out.lo.ctypes.data, out.hi.ctypes.data,
                                            out.flags.ctypes.data, insert.ctypes.data, st.ctypes.data), "tbo")
        return insert, st

    def tbo_device(self, d_bases, d_quals, d_offsets, n_reads, max_read_len, d_lo, d_hi, d_flags, d_insert=None, d_stats=None,
                   stream=None, cfg=None):
        cfg = cfg or self.tbo_cfg()

        def ptr(x):
            return None if x is None else (x.data_ptr() if hasattr(x, "data_ptr") else int(x))
        self._check(self.lib.bbduk_b200_tbo_device(self.h, C.byref(cfg), ptr(d_bases), ptr(d_quals), ptr(d_offsets), n_reads,
                                                   max_read_len, ptr(d_lo), ptr(d_hi), ptr(d_flags), ptr(d_insert), ptr(d_stats),
                                                   ptr(stream)), "tbo_device")

    # -- quality trimming + quality / length / N filters (jgi/BBDuk.java:3074-3170) ---------------------
    def qtrim_cfg(self, **kw):
        from ._abi import BBDukQtrimCfg
        c = BBDukQtrimCfg()
        self.lib.bbduk_b200_qtrim_cfg_default(C.byref(c))
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def qtrim(self, bases, quals, offsets, paired, out, cfg):
        """HOST buffers; `out` is the Outputs of process() (after tbo): lo / hi / flags are updated in place.
        -> [readsQTrimmed, basesQTrimmed, readsQFiltered, basesQFiltered, readsNFiltered, basesNFiltered, readsPolyTrimmed,
        basesPolyTrimmed]"""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        q = None if quals is None else np.ascontiguousarray(quals, np.uint8)
        st = np.zeros(8, np.int64)
        self._check(self.lib.bbduk_b200_qtrim(self.h, C.byref(cfg), bases.ctypes.data, None if q is None else q.ctypes.data,
                                              offsets.ctypes.data, len(offsets) - 1, int(bool(paired)), out.lo.ctypes.data,
                                              out.hi.ctypes.data, out.flags.ctypes.data, st.ctypes.data), "qtrim")
        return st

    def qtrim_device(self, d_bases, d_quals, d_offsets, n_reads, paired, d_lo, d_hi, d_flags, cfg, d_stats=None, stream=None):
        def ptr(x):
            if x is None:
                return None
            return x.data_ptr() if hasattr(x, "data_ptr") else int(x)
        self._check(self.lib.bbduk_b200_qtrim_device(self.h, C.byref(cfg), ptr(d_bases), ptr(d_quals), ptr(d_offsets), n_reads,
                                                     int(bool(paired)), ptr(d_lo), ptr(d_hi), ptr(d_flags), ptr(d_stats),
                                                     ptr(stream)), "qtrim_device")

    # -- low-entropy read filter (jgi/BBDuk.java:3175-3186, tracker/EntropyTracker.java) -----------------
    def entropy_cfg(self, **kw):
        from ._abi import BBDukEntropyCfg
        c = BBDukEntropyCfg()
        self.lib.bbduk_b200_entropy_cfg_default(C.byref(c))
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def entropy(self, bases, offsets, paired, out, cfg):
        """HOST buffers; `out` is the Outputs of process() (after tbo / qtrim): hi / flags are updated in place.
        -> [readsEFiltered, basesEFiltered]"""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        st = np.zeros(2, np.int64)
        self._check(self.lib.bbduk_b200_entropy(self.h, C.byref(cfg), bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                                int(bool(paired)), out.lo.ctypes.data, out.hi.ctypes.data, out.flags.ctypes.data,
                                                st.ctypes.data), "entropy")
        return st

    def entropy_mask(self, bases, offsets, paired, out, cfg, mode):
        """entropymask (mode 1: to N, 2: to lower case) / entropytrim (mode 3) on HOST buffers; `out` is the Outputs so far: mode 3
        updates out.lo / out.hi. -> (mask words, mask_off, [readsEFiltered, basesEFiltered]); bit j of a read's words = base j of
        its kept interval"""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        words = (np.diff(offsets) + 31) // 32
        mask_off = np.zeros(len(offsets), np.int64)
        np.cumsum(words, out=mask_off[1:])
        bits = np.zeros(max(1, int(mask_off[-1])), np.uint32)
        st = np.zeros(2, np.int64)
        self._check(self.lib.bbduk_b200_entropy_mask(self.h, C.byref(cfg), int(mode), bases.ctypes.data, offsets.ctypes.data,
                                                     len(offsets) - 1, int(bool(paired)), out.lo.ctypes.data, out.hi.ctypes.data,
                                                     out.flags.ctypes.data, bits.ctypes.data, mask_off.ctypes.data, st.ctypes.data),
                    "entropy_mask")
        return bits, mask_off, st

    def entropy_device(self, d_bases, d_offsets, n_reads, paired, d_lo, d_hi, d_flags, cfg, d_stats=None, stream=None):
        def ptr(x):
            if x is None:
                return None
            return x.data_ptr() if hasattr(x, "data_ptr") else int(x)
        self._check(self.lib.bbduk_b200_entropy_device(self.h, C.byref(cfg), ptr(d_bases), ptr(d_offsets), n_reads, int(bool(paired)),
                                                       ptr(d_lo), ptr(d_hi), ptr(d_flags), ptr(d_stats), ptr(stream)), "entropy_device")

    # -- the whole device part of the per-pair loop in one call -------------------------------------------
    def process_chain(self, bases, quals, offsets, paired, tbo=None, qtrim=None, entropy=None, out=None):
        """HOST buffers: k-mer block, then the given steps (their cfg structs, or None to skip) with ONE upload of the batch.
        out: caller-owned Outputs (e.g. pinned arrays; lo, hi and flags are required, arrays set to None are not downloaded).
        -> (Outputs, BBDukStats, tbo stats2, qtrim stats8, entropy stats2)"""
        from ._abi import BBDukChainCfg
        c = BBDukChainCfg()
        self.lib.bbduk_b200_chain_cfg_default(C.byref(c))
        for name, step in (("tbo", tbo), ("qtrim", qtrim), ("entropy", entropy)):
            if step is not None:
                setattr(c, "do_" + name, 1)
                setattr(c, name, step)
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        q = None if quals is None else np.ascontiguousarray(quals, np.uint8)
        n = len(offsets) - 1
        if out is None:
            out = Outputs(n)
        o = out.struct()
        st = BBDukStats()
        t2, q8, e2 = np.zeros(2, np.int64), np.zeros(8, np.int64), np.zeros(2, np.int64)
        self._check(self.lib.bbduk_b200_process_chain(self.h, C.byref(c), bases.ctypes.data, None if q is None else q.ctypes.data,
                                                      offsets.ctypes.data, n, int(bool(paired)), C.byref(o), C.byref(st),
                                                      t2.ctypes.data, q8.ctypes.data, e2.ctypes.data), "process_chain")
        return out, st, t2, q8, e2

    def set_max_read_len(self, n):
        self._check(self.lib.bbduk_b200_set_max_read_len(self.h, int(n)), "set_max_read_len")

    def scaffold_counts(self):
        n = self.n_scaffolds + 1
        rc_ = np.zeros(n, np.int64)
        bc = np.zeros(n, np.int64)
        self._check(self.lib.bbduk_b200_scaffold_counts(self.h, rc_.ctypes.data, bc.ctypes.data, n), "scaffold_counts")
        return rc_, bc

    def transfer_bytes(self):
        """(host->device, device->host) bytes process() / process_packed() have moved so far"""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.bbduk_b200_transfer_bytes(self.h, C.byref(a), C.byref(b)), "transfer_bytes")
        return a.value, b.value

    @property
    def launches(self):
        return int(self.lib.bbduk_b200_launch_count(self.h))


class BBDuk:
    """reads-in / reads-out tool: `BBDuk(["in=r.fq", "ref=adapters.fa", "ktrim=r", "k=23", ...]).process()`.

    FASTQ handling is deliberately simple host code; the k-mer block of every batch runs on the GPU.
    Output records follow the reference's routing (jgi/BBDuk.java:3190-3254): pairs that are not
    removed go to out (trimmed), removed pairs go to outm when given."""

    def __init__(self, args, generation=GEN_JGI, resources=None):
        self.cfg, self.io = parse_args(args, generation)
        self.resources = resources
        self.index = BBDukIndexGPU(self.cfg)
        self.scaffold_names = [""]
        for ref in self.io["ref"]:
            path = ref
            if ref == "adapters" and resources:
                path = os.path.join(resources, "adapters.fa")
            names, b, off = read_fasta(path)
            self.scaffold_names += names
            self.index.add_ref(b, off)
        if self.io["literal"]:
            b, off = pack([s.encode() for s in self.io["literal"]])
            self.scaffold_names += [str(len(self.scaffold_names) + i) for i in range(len(self.io["literal"]))]
            self.index.add_ref(b, off)
        self.stored_kmers = self.index.finalize()
        self.stats = None
        self.tbo_stats = None  # [readsTrimmedByOverlap, basesTrimmedByOverlap]
        self.entropy_stats = None  # [readsEFiltered, basesEFiltered]
        self.qtrim_stats = None  # [reads/bases QTrimmed, reads/bases QFiltered, reads/bases NFiltered, reads/bases PolyTrimmed]

    def process_arrays(self, bases, offsets, paired):
        want_mask = bool(self.cfg.ktrim_n)
        return self.index.process(bases, offsets, paired, want_mask=want_mask)

    def process_native(self, block_bytes=None):
        """FASTQ in / FASTQ out through the native feed (include/fastq_b200.h); ktrim / kfilter modes.
        The input is streamed in blocks of `block_bytes` of text per file (default 256 MiB, BBDUK_B200_FEED_BLOCK overrides):
        whole records only, the same number from both mate files; outputs are appended block by block, counters summed."""
        import os

        from ._abi import BBDukStats
        from .fastq import FastqBatch, iter_fastq_blocks
        io = self.io
        paired = bool(io["in2"] or io["interleaved"])
        per = 2 if paired else 1
        want_tbo, want_q, want_e = bool(io["tbo"] and paired), self._wants_qtrim(), io["entropy"] >= 0
        block = int(block_bytes or os.environ.get("BBDUK_B200_FEED_BLOCK", 256 << 20))
        total = BBDukStats()
        sums = {"tbo": np.zeros(2, np.int64), "q": np.zeros(8, np.int64), "e": np.zeros(2, np.int64)}
        routes = [(removed, path, sel) for removed, p1, p2 in ((False, io["out1"], io["out2"]), (True, io["outm1"], io["outm2"])) if p1
                  for path, sel in (((p1, 1), (p2, 2)) if p2 else ((p1, 0),))]
        files = {path: open(path, "wb") for _, path, _ in routes}
        try:
            for text1, text2 in iter_fastq_blocks(io["in1"], io["in2"] or None, block, unit=per):
                fb = FastqBatch(text1, text2)
                bases, offsets = fb.arrays()
                if want_tbo or want_q or want_e:
                    # one upload of the batch, the steps hand lo / hi / flags to each other on the device
                    quals = fb.quals() if (want_tbo or want_q) else None
                    out, st, t2, q8, e2 = self.index.process_chain(bases, quals, offsets, paired, tbo=self._tbo_cfg() if want_tbo else None,
                                                                   qtrim=self._qtrim_cfg() if want_q else None,
                                                                   entropy=self._entropy_cfg() if want_e else None)
                    for key, v in (("tbo", t2), ("q", q8), ("e", e2)):
                        if v is not None:
                            sums[key] += np.asarray(v, np.int64)
                else:
                    out, st = self.process_arrays(bases, offsets, paired)
                for name, _ in BBDukStats._fields_:
                    setattr(total, name, getattr(total, name) + getattr(st, name))
                for removed, path, sel in routes:
                    fb.format(per, out.lo, out.hi, out.flags, removed=removed, mate_sel=sel, trim_removed=bool(io["ottm"])).tofile(files[path])
        finally:
            for f in files.values():
                f.close()
        self.tbo_stats = sums["tbo"] if want_tbo else None
        self.qtrim_stats = sums["q"] if want_q else None
        self.entropy_stats = sums["e"] if want_e else None
        self.stats = total
        self._write_stats()
        return total

    def _tbo_cfg(self):
        io = self.io
        return self.index.tbo_cfg(strict_overlap=int(io["strictoverlap"]), min_overlap=io["minoverlap"], min_insert=io["mininsert"])

    def _tbo(self, bases, quals, offsets, out):
        """the tbo block (jgi/BBDuk.java:2878-2926) on the batch the k-mer block just answered; updates out.hi / out.flags"""
        _, st = self.index.tbo(bases, quals, offsets, out, self._tbo_cfg())
        return st

    def _wants_qtrim(self):
        io = self.io
        poly = any(io[k] > 0 for k in ("trimpolya", "trimpolygleft", "trimpolygright", "filterpolyg", "trimpolycleft",
                                       "trimpolycright", "filterpolyc"))
        return bool(io["qtrim_left"] or io["qtrim_right"] or io["mbq"] > 0 or io["maxns"] >= 0 or io["maxlen"] > 0 or io["tbo"] or poly or
                    io["maq"] > 0 or io["maxnrate"] < 1 or io["mcb"] > 0 or io["minbasefrequency"] > 0)

    def _qtrim(self, bases, quals, offsets, paired, out):
        """quality trimming, minlen / maxlen, mbq, maxns (jgi/BBDuk.java:3074-3170); updates out.lo / out.hi / out.flags"""
        return self.index.qtrim(bases, quals, offsets, paired, out, self._qtrim_cfg())

    def _qtrim_cfg(self):
        io = self.io
        return self.index.qtrim_cfg(qtrim_left=int(io["qtrim_left"]), qtrim_right=int(io["qtrim_right"]), trimq=io["trimq"],
                                   min_base_quality=io["mbq"], max_ns=io["maxns"], max_read_length=io["maxlen"],
                                   trim_poly_a=io["trimpolya"], trim_poly_g_left=io["trimpolygleft"],
                                   trim_poly_g_right=io["trimpolygright"], filter_poly_g=io["filterpolyg"],
                                   trim_poly_c_left=io["trimpolycleft"], trim_poly_c_right=io["trimpolycright"],
                                   filter_poly_c=io["filterpolyc"], max_non_poly=io["maxnonpoly"], min_avg_quality=io["maq"],
                                   min_avg_quality_bases=io["maqb"], max_n_rate=io["maxnrate"], min_consecutive_bases=io["mcb"],
                                   min_base_frequency=io["minbasefrequency"], trim_mode=io["trim_mode"],
                                   window_length=io["window_length"], min_good_interval=io["min_good_interval"])

    def _entropy_cfg(self):
        io = self.io
        return self.index.entropy_cfg(cutoff=io["entropy"], k=io["entropyk"], window=io["entropywindow"])

    def _entropy(self, bases, offsets, paired, out):
        """the low-entropy read filter (jgi/BBDuk.java:3175-3186); updates out.flags"""
        return self.index.entropy(bases, offsets, paired, out, self._entropy_cfg())

    def _write_stats(self):
        io = self.io
        if io["stats"]:
            rc_, bc = self.index.scaffold_counts()
            with open(io["stats"], "w") as f:
                f.write("#Name\tReads\tBases\n")
                for i in np.argsort(-rc_, kind="stable"):
                    if i > 0 and rc_[i] > 0:
                        f.write(f"{self.scaffold_names[i]}\t{rc_[i]}\t{bc[i]}\n")

    def process(self, native=True):
        io = self.io
        e_mode = 3 if io["entropytrim"] else io["entropymask"]  # entropy masking / trimming rewrites bases: plain-Python feed
        if e_mode:
            if io["entropy"] < 0:
                raise ValueError("Entropy masking/trimming operations require the entropy flag to be set.")  # jgi/BBDuk.java:1028
            if any(io[x] for x in ("trimpolya", "trimpolygleft", "trimpolygright", "filterpolyg", "trimpolycleft", "trimpolycright", "filterpolyc")):
                raise NotImplementedError("entropy masking / trimming together with poly-X steps (they sit on either side of it in one launch)")
        if native and not self.cfg.ktrim_n and not self.cfg.ksplit and not e_mode:
            return self.process_native()
        n1, s1, q1 = read_fastq(io["in1"])
        paired = False
        if io["in2"]:
            n2, s2, q2 = read_fastq(io["in2"])
            paired = True
            names = [x for p in zip(n1, n2) for x in p]
            seqs = [x for p in zip(s1, s2) for x in p]
            quals = [x for p in zip(q1, q2) for x in p]
        else:
            names, seqs, quals = n1, s1, q1
            if io["interleaved"]:
                paired = True
        bases, offsets = pack(seqs)
        out, st = self.process_arrays(bases, offsets, paired)
        self.stats = st
        if io["tbo"] and paired:
            self.tbo_stats = self._tbo(bases, pack(quals)[0], offsets, out)
        if e_mode:  # jgi/BBDuk.java:3055-3067: before quality trimming; the later entropy FILTER is then off (:3175)
            bits, moff, self.entropy_stats = self.index.entropy_mask(bases, offsets, paired, out, self._entropy_cfg(), e_mode)
            if e_mode != 3:  # maskFromBitset :4505-4526: an N stays an N (and keeps its quality), lower case stays as it is
                qarr = pack(quals)[0]
                for i in range(len(seqs)):
                    a = int(offsets[i]) + int(out.lo[i])
                    for j in range(int(out.hi[i]) - int(out.lo[i])):
                        if (int(bits[moff[i] + (j >> 5)]) >> (j & 31)) & 1:
                            if e_mode == 1 and bases[a + j] != ord("N"):
                                bases[a + j] = ord("N")
                                qarr[a + j] = 33
                            elif e_mode == 2:
                                bases[a + j] = ord(chr(bases[a + j]).lower())
                seqs = [bytes(bases[offsets[i]:offsets[i + 1]]) for i in range(len(seqs))]
                quals = [bytes(qarr[offsets[i]:offsets[i + 1]]) for i in range(len(seqs))]
        if self._wants_qtrim():
            self.qtrim_stats = self._qtrim(bases, pack(quals)[0], offsets, paired, out)
        if io["entropy"] >= 0 and not e_mode:
            self.entropy_stats = self._entropy(bases, offsets, paired, out)
        sinks = {}

        def sink(path):
            if path and path not in sinks:
                sinks[path] = open(path, "wb")
            return sinks.get(path)

        per = 2 if paired else 1
        sym = bytes([self.cfg.trim_symbol])
        for u in range(len(seqs) // per):
            removed = bool(out.flags[u * per] & F_REMOVED)
            for q in range(per):
                i = u * per + q
                dest = (io["outm2"] if q and io["outm2"] else io["outm1"]) if removed else \
                       (io["out2"] if q and io["out2"] else io["out1"])
                f = sink(dest)
                if f is None:
                    continue
                s, ql = bytearray(seqs[i]), bytearray(quals[i])
                if out.maskbits is not None:
                    w0 = int(out.mask_off[i])
                    for j in range(len(s)):
                        if (int(out.maskbits[w0 + (j >> 5)]) >> (j & 31)) & 1:
                            if self.cfg.kmask_lowercase:
                                s[j:j + 1] = bytes(s[j:j + 1]).lower()
                            else:
                                s[j:j + 1] = sym
                                if sym == b"N":
                                    ql[j] = 33
                lo, hi = int(out.lo[i]), int(out.hi[i])
                if removed and not io["ottm"]:
                    lo, hi = 0, len(s)
                f.write(b"@" + names[i] + b"\n" + bytes(s[lo:hi]) + b"\n+\n" + bytes(ql[lo:hi]) + b"\n")
                if out.flags[i] & F_SPLIT:
                    a = int(out.count[i])
                    f.write(b"@" + names[i] + b"\n" + bytes(s[a:len(s) - 1]) + b"\n+\n" + bytes(ql[a:len(s) - 1]) + b"\n")
        for f in sinks.values():
            f.close()
        self._write_stats()
        return st
