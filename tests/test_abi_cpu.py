"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
derives the same constants as the oracle, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import itertools
import os
import re

import numpy as np
import pytest

from bbtools_b200 import _lib, make_cfg
from bbtools_b200._abi import BBDukCfg
from bbtools_b200.bbduk import parse_args
from oracle.oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bbduk_b200.h")).read()
    names = set(re.findall(r"BBDUK_API[^;(]*?\b(bbduk_b200_\w+)\s*\(", text))
    text = open(os.path.join(ROOT, "include", "kcount_b200.h")).read()
    names |= set(re.findall(r"KCOUNT_API[^;(]*?\b(kcount_b200_\w+)\s*\(", text))
    text = open(os.path.join(ROOT, "include", "seal_b200.h")).read()
    names |= set(re.findall(r"SEAL_API[^;(]*?\b(seal_b200_\w+)\s*\(", text))
    text = open(os.path.join(ROOT, "include", "fastq_b200.h")).read()
    names |= set(re.findall(r"FASTQ_API[^;(]*?\b(fastq_b200_\w+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported"
    assert sorted(s[0] for s in _lib.SYMBOLS) == names  # the binding covers the header exactly
    from bbtools_b200._abi import ABI_VERSION
    assert lib.bbduk_b200_version() == ABI_VERSION == 2


def test_cfg_struct_layout_matches_header():
    lib = _lib.load()
    c = BBDukCfg()
    lib.bbduk_b200_cfg_default(C.byref(c))
    assert c.struct_size == C.sizeof(BBDukCfg)
    ref = make_cfg()
    for name, _ in BBDukCfg._fields_:
        if name != "reserved":
            assert getattr(c, name) == getattr(ref, name), name


def describe(cfg):
    lib = _lib.load()
    v = np.zeros(16, np.int64)
    rc = lib.bbduk_b200_describe_cfg(C.byref(cfg), v.ctypes.data)
    if rc:
        raise ValueError(lib.bbduk_b200_last_error(None).decode())
    names = ["k", "kbig", "mink", "useShortKmers", "maskMiddle", "midMaskLen", "minlen", "minlen2", "minminlen",
             "forbidNs", "hdist", "hdist2", "middleMask", "mask", "kfilter", "rieb"]
    return dict(zip(names, (int(x) for x in v)))


def test_derived_constants_match_oracle():
    ks = [1, 5, 11, 12, 23, 27, 31, 32, 40, 0]
    minks = [-1, 4, 11, 31]
    modes = [dict(), dict(ktrim_right=1), dict(ktrim_left=1, ktrim_right=1), dict(ktrim_n=1), dict(ksplit=1)]
    n = 0
    for k, mink, mode, hd, mm, gen in itertools.product(ks, minks, modes, (0, 1, 2), (0, 1), (0, 1)):
        for extra in (dict(), dict(mid_mask_len=3), dict(edist=1), dict(hdist2=0, forbid_ns=1, require_both_bad=1)):
            cfg = make_cfg(k=k, mink=mink, hdist=hd, mask_middle=mm, generation=gen, **mode, **extra)
            try:
                want = Oracle(cfg).derived()
            except ValueError:
                with pytest.raises(ValueError):
                    describe(cfg)
                continue
            if gen == 1 and mink == -1 and mode and want["useShortKmers"]:
                continue
            got = describe(cfg)
            assert got == want, (k, mink, mode, hd, mm, gen, extra)
            n += 1
    assert n > 2000


def test_parse_args_mirrors_reference_flags():
    cfg, io = parse_args("in=a.fq in2=b.fq out=c.fq ref=adapters ktrim=r k=23 mink=11 hdist=1 tpe".split())
    assert (cfg.ktrim_right, cfg.ktrim_left, cfg.k, cfg.mink, cfg.hdist, cfg.trim_pairs_evenly) == (1, 0, 23, 11, 1, 1)
    assert io["ref"] == ["adapters"] and io["in2"] == "b.fq"
    cfg, _ = parse_args(["ktrim=rl", "mm=f", "mkh=3", "rieb=f", "minlen=25", "mlf=0.5"])
    assert cfg.ktrim_left and cfg.ktrim_right and not cfg.mask_middle and cfg.max_bad_kmers == 2
    assert cfg.require_both_bad == 1 and cfg.min_read_length == 25 and abs(cfg.min_len_fraction - 0.5) < 1e-7
    cfg, _ = parse_args(["kmask=lc"])
    assert cfg.ktrim_n and cfg.kmask_lowercase
    cfg, _ = parse_args(["ktrim=X"])
    assert cfg.ktrim_n and cfg.trim_symbol == ord("X")
    cfg, _ = parse_args(["ktrimtips=40"])
    assert cfg.ktrim_left and cfg.ktrim_right and cfg.restrict_left == 40 and cfg.restrict_right == 40
    cfg, _ = parse_args(["mm=3"])
    assert cfg.mid_mask_len == 3 and cfg.mask_middle
    with pytest.raises(ValueError):
        parse_args(["nosuchflag=1"])


def test_no_cpu_fallback():
    """Without a CUDA device create() must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    cfg = make_cfg(k=23, ktrim_right=1)
    rc = lib.bbduk_b200_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.bbduk_b200_last_error(None)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bbtools_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f in ("synth.py",), f"{f} mentions the oracle"


def test_host_packer_matches_the_codec_tables():
    """bbduk_b200_pack_bases (AVX2 or scalar) against a table-driven numpy restatement, every byte value included"""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    code = np.zeros(256, np.uint32)
    valid = np.zeros(256, np.uint32)
    for i, ch in enumerate("ACGT"):
        code[ord(ch)] = code[ord(ch.lower())] = i
        valid[ord(ch)] = valid[ord(ch.lower())] = 1
    code[ord("U")] = code[ord("u")] = 3
    valid[ord("U")] = valid[ord("u")] = 1
    for n in (0, 1, 15, 16, 17, 31, 32, 33, 1000, 4099, 100003):
        b = rng.integers(0, 256, n).astype(np.uint8)
        if n > 50:
            b[10:n // 2] = np.frombuffer(b"ACGTacgtUuNn", np.uint8)[rng.integers(0, 12, n // 2 - 10)]
        g = (n + 15) // 16
        F = np.full(g + 1, 0xDEADBEEF, np.uint32)
        D = np.full(g + 1, 0xBEEF, np.uint16)
        assert lib.bbduk_b200_pack_bases(b.ctypes.data, n, F.ctypes.data, D.ctypes.data) == 0
        pad = np.zeros(g * 16, np.uint8)
        pad[:n] = b
        inside = (np.arange(g * 16) < n).astype(np.uint32)
        c = (code[pad] * inside).reshape(g, 16)
        v = (valid[pad] * inside).reshape(g, 16)
        wantF = (c << (30 - 2 * np.arange(16, dtype=np.uint32))[None, :]).sum(axis=1).astype(np.uint32)
        wantD = (v << (15 - np.arange(16, dtype=np.uint32))[None, :]).sum(axis=1).astype(np.uint16)
        assert np.array_equal(F[:g], wantF), n
        assert np.array_equal(D[:g], wantD), n
        assert F[g] == 0xDEADBEEF and D[g] == 0xBEEF  # nothing written past the end


@pytest.mark.parametrize("shim", ["BBDukCuda.c", "KCountCuda.c", "SealCuda.c"])
def test_jni_shims_compile_and_link_against_the_c_abi(shim, tmp_path):
    """No JDK in the image: the shims are compiled with a stub jni.h (tests/stubs) and linked against the library,
    so that every C-ABI call they make is checked against the real prototypes and resolves."""
    import subprocess
    _lib.load()
    out = tmp_path / (shim + ".so")
    cmd = ["gcc", "-O2", "-std=c99", "-Wall", "-Werror", "-fPIC", "-shared", "-Wl,--no-undefined",
           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", shim),
           "-L", os.path.join(ROOT, "bbtools_b200"), "-lbbduk_b200", "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    syms = subprocess.run(["nm", "-D", "--defined-only", str(out)], capture_output=True, text=True).stdout
    assert "Java_" in syms
