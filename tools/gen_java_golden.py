#!/usr/bin/env python
"""Extract the abstract methods of the reference's plugin seam, bbduk/BBDukIndex.java, into tests/golden/bbdukindex_abstract.json
(tests/test_java_shim_cpu.py checks java/bbduk/BBDukIndexGPU.java against it; /root/reference does not exist on the GPU box).

    python tools/gen_java_golden.py [/root/reference]
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def java_methods(text, want_abstract):
    """[(return type, name, [param types])] of the method declarations of one Java class body"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = []
    pat = re.compile(r"^\s*((?:@Override\s+)?(?:(?:public|private|protected|static|final|synchronized|abstract|native)\s+)*)"
                     r"([\w\[\]<>\.]+)\s+(\w+)\s*\(([^)]*)\)\s*(?:throws [\w\., ]+)?\s*([;{])", re.M)
    for m in pat.finditer(text):
        mods, ret, name, params, end = m.groups()
        if ret in ("return", "new", "else", "throw"):
            continue
        is_abs = "abstract" in mods.split()
        if want_abstract != is_abs:
            continue
        ptypes = []
        for prm in [x.strip() for x in params.split(",") if x.strip()]:
            toks = [t for t in prm.split() if t != "final"]
            ptypes.append(toks[0])
        out.append((ret, name, ptypes, "native" in mods.split()))
    return out


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = open(os.path.join(ref, "current", "bbduk", "BBDukIndex.java")).read()
    meths = [{"returns": r, "name": n, "params": p} for r, n, p, _ in java_methods(src, True)]
    dst = os.path.join(ROOT, "tests", "golden", "bbdukindex_abstract.json")
    with open(dst, "w") as f:
        json.dump({"source": "current/bbduk/BBDukIndex.java", "abstract_methods": meths}, f, indent=1)
    print(len(meths), "abstract methods ->", dst)


if __name__ == "__main__":
    main()
