"""The Java side of the drop-in boundary cannot be compiled here (no JDK in the image), so it is checked by inspection, mechanically:
java/bbduk/BBDukIndexGPU.java must implement every abstract method of the reference's plugin seam (bbduk/BBDukIndex.java:28-99,
kept as tests/golden/bbdukindex_abstract.json), every `native` method must have its JNI entry point in jni/BBDukCuda.c with the
matching argument count, marshal() must fill bbduk_cfg up to the field the header says, and the four patches to the reference
(java/patches/*.diff) must apply cleanly to the reference's files where those are present."""
import json
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_java_golden import java_methods  # noqa: E402

REF = "/root/reference"
GPU_JAVA = os.path.join(ROOT, "java", "bbduk", "BBDukIndexGPU.java")


def golden():
    with open(os.path.join(ROOT, "tests", "golden", "bbdukindex_abstract.json")) as f:
        return json.load(f)["abstract_methods"]


def test_golden_list_is_current():
    if not os.path.exists(os.path.join(REF, "current", "bbduk", "BBDukIndex.java")):
        pytest.skip("reference tree not present")
    src = open(os.path.join(REF, "current", "bbduk", "BBDukIndex.java")).read()
    now = [{"returns": r, "name": n, "params": p} for r, n, p, _ in java_methods(src, True)]
    assert now == golden() and len(now) == 16


def test_every_abstract_method_is_implemented():
    src = open(GPU_JAVA).read()
    assert re.search(r"public final class BBDukIndexGPU extends BBDukIndex\b", src)
    assert "super(" not in src, "bbduk/BBDukIndex.java has only the implicit constructor"
    have = {(n, tuple(p)): r for r, n, p, _ in java_methods(src, False)}
    for m in golden():
        key = (m["name"], tuple(m["params"]))
        assert key in have, f"BBDukIndexGPU lacks {m['returns']} {m['name']}({', '.join(m['params'])})"
        assert have[key] == m["returns"], f"{m['name']}: returns {have[key]}, the seam says {m['returns']}"


def test_native_methods_have_jni_entry_points():
    src = open(GPU_JAVA).read()
    c = open(os.path.join(ROOT, "jni", "BBDukCuda.c")).read()
    natives = [(n, p) for _, n, p, nat in java_methods(src, False) if nat]
    assert len(natives) >= 14
    for name, params in natives:
        m = re.search(r"Java_bbduk_BBDukIndexGPU_%s\(JNIEnv \*env, jclass cls([^)]*)\)" % name, c)
        assert m, f"jni/BBDukCuda.c lacks Java_bbduk_BBDukIndexGPU_{name}"
        n_c = len([x for x in m.group(1).split(",") if x.strip()])
        assert n_c == len(params), f"{name}: Java declares {len(params)} arguments, the shim takes {n_c}"
    assert "GetPrimitiveArrayCritical" not in re.sub(r"/\*.*?\*/", "", c, flags=re.S), \
        "no critical region may be held across CUDA work (ADVICE r1)"


def test_marshal_covers_the_cfg_struct():
    """marshal() sends bbduk_cfg in declaration order: count the 32-bit fields of the header up to and including minlen2"""
    hdr = open(os.path.join(ROOT, "include", "bbduk_b200.h")).read()
    body = hdr[hdr.index("typedef struct bbduk_cfg {"):hdr.index("} bbduk_cfg;")]
    fields = re.findall(r"^\s*(?:int32_t|float)\s+(\w+)(?:\[\d+\])?;", body, flags=re.M)
    upto = fields.index("minlen2") + 1
    src = open(GPU_JAVA).read()
    arr = src[src.index("return new int[] {", src.index("private static int[] marshal")):]
    arr = arr[:arr.index("};")]
    arr = re.sub(r"/\*.*?\*/", "", arr, flags=re.S)
    depth, n = 0, 1
    for ch in arr[arr.index("{") + 1:]:
        depth += ch in "([" 
        depth -= ch in ")]"
        n += (ch == "," and depth == 0)
    assert n == upto, f"marshal() sends {n} ints, bbduk_cfg has {upto} fields up to minlen2: {fields[:upto]}"


@pytest.mark.parametrize("name", ["BBDukParser", "BBDukLoader", "BBDukProcessorS", "BBDukS"])
def test_patches_apply_to_the_reference(name, tmp_path):
    ref_file = os.path.join(REF, "current", "bbduk", name + ".java")
    if not os.path.exists(ref_file) or not shutil.which("patch"):
        pytest.skip("reference tree (or patch) not present")
    d = tmp_path / "current" / "bbduk"
    d.mkdir(parents=True)
    shutil.copy(ref_file, d / (name + ".java"))
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(ROOT, "java", "patches", name + ".diff")],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_patches_only_use_what_the_java_files_offer():
    """every BBDukIndexGPU / BBDukGpuBatch member the patches call exists in our Java files"""
    gpu = open(GPU_JAVA).read() + open(os.path.join(ROOT, "java", "bbduk", "BBDukGpuBatch.java")).read()
    used = set()
    for name in ("BBDukLoader", "BBDukProcessorS", "BBDukS"):
        txt = open(os.path.join(ROOT, "java", "patches", name + ".diff")).read()
        added = "\n".join(ln[1:] for ln in txt.splitlines() if ln.startswith("+") and not ln.startswith("+++"))
        used |= set(re.findall(r"(?:\bgpu|\bgpuBatch|\(\(BBDukIndexGPU\)index\))\.(\w+)", added))
    assert used >= {"addScaffolds", "finalizeTable", "refKmersSeen", "run", "apply", "fetchScaffoldCounts"}
    for member in used:
        assert re.search(r"\b%s\b\s*[(;=,]" % member, gpu), f"patches use .{member} which the Java files do not declare"


# ---- Seal (java/jgi/SealGPU.java, jni/SealCuda.c, java/patches/Seal.diff) ------------------------------------------------
SEAL_JAVA = os.path.join(ROOT, "java", "jgi", "SealGPU.java")


def test_seal_native_methods_have_jni_entry_points():
    src = open(SEAL_JAVA).read()
    c = open(os.path.join(ROOT, "jni", "SealCuda.c")).read()
    natives = [(n, p) for _, n, p, nat in java_methods(src, False) if nat]
    assert sorted(n for n, _ in natives) == ["addRefNative", "createNative", "destroyNative", "finalizeNative", "lastErrorNative",
                                             "processNative", "scaffoldCountsNative"]
    for name, params in natives:
        m = re.search(r"Java_jgi_SealGPU_%s\(JNIEnv \*env, jclass cls([^)]*)\)" % name, c)
        assert m, f"jni/SealCuda.c lacks Java_jgi_SealGPU_{name}"
        n_c = len([x for x in m.group(1).split(",") if x.strip()])
        assert n_c == len(params), f"{name}: Java declares {len(params)} arguments, the shim takes {n_c}"
    assert len(re.findall(r"JNIEXPORT", c)) == len(natives)  # and the shim exports nothing the class does not declare
    assert "GetPrimitiveArrayCritical" not in re.sub(r"/\*.*?\*/", "", c, flags=re.S)
    # createNative's int[] carries the 17 values the shim reads, in the header's order of the integer fields
    m = re.search(r"final int\[\] cfg=\{(.*?)\};", src, flags=re.S)
    assert m and len([x for x in m.group(1).split(",") if x.strip()]) == 17


def test_seal_patch_applies_to_the_reference(tmp_path):
    ref_file = os.path.join(REF, "current", "jgi", "Seal.java")
    if not os.path.exists(ref_file) or not shutil.which("patch"):
        pytest.skip("reference tree (or patch) not present")
    d = tmp_path / "current" / "jgi"
    d.mkdir(parents=True)
    shutil.copy(ref_file, d / "Seal.java")
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(ROOT, "java", "patches", "Seal.diff")],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_seal_patch_only_uses_what_sealgpu_offers():
    gpu = open(SEAL_JAVA).read()
    txt = open(os.path.join(ROOT, "java", "patches", "Seal.diff")).read()
    added = "\n".join(ln[1:] for ln in txt.splitlines() if ln.startswith("+") and not ln.startswith("+++"))
    used = set(re.findall(r"(?:\bgpu|\bgpuHits|\bSealGPU)\.(\w+)", added))
    assert used >= {"createIfServed", "addRef", "finalizeTable", "refKmers", "match", "addScaffoldCounts", "close", "sites", "assigned",
                    "removed", "readsMatched", "basesMatched", "readsUnmatched", "basesUnmatched", "Result"}
    for member in used:
        assert re.search(r"\b%s\b\s*[(;=,{]" % member, gpu), f"Seal.diff uses .{member} which SealGPU.java does not declare"
    # the call of createIfServed passes as many arguments as the method declares
    call = re.search(r"SealGPU\.createIfServed\((.*?)\);", added, flags=re.S).group(1)
    decl = re.search(r"public static SealGPU createIfServed\((.*?)\)\{", gpu, flags=re.S).group(1)
    assert len(call.split(",")) == len(decl.split(",")) == 22
    # every field of Seal the patch reads exists in the reference (when it is present)
    ref_file = os.path.join(REF, "current", "jgi", "Seal.java")
    if os.path.exists(ref_file):
        ref = open(ref_file).read()
        for name in re.findall(r"[!( ]([a-zA-Z_]\w*)(?:<=0|<0|==null| &&|\))", re.search(r"gpuPreambleIsIdle\(\)\{(.*?)\}", added, flags=re.S).group(1)):
            if name not in ("return", "null"):
                assert re.search(r"\b%s\b" % name, ref), name


def test_kcount_native_methods_have_jni_entry_points():
    src = open(os.path.join(ROOT, "java", "kmer", "KmerTableSetGPU.java")).read()
    c = open(os.path.join(ROOT, "jni", "KCountCuda.c")).read()
    natives = [(n, p) for _, n, p, nat in java_methods(src, False) if nat]
    assert len(natives) == len(re.findall(r"JNIEXPORT", c)) == 6
    for name, params in natives:
        m = re.search(r"Java_kmer_KmerTableSetGPU_%s\(JNIEnv \*env, jclass cls([^)]*)\)" % name, c)
        assert m, f"jni/KCountCuda.c lacks Java_kmer_KmerTableSetGPU_{name}"
        assert len([x for x in m.group(1).split(",") if x.strip()]) == len(params), name
    assert "GetPrimitiveArrayCritical" not in re.sub(r"/\*.*?\*/", "", c, flags=re.S)
