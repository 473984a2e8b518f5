#!/usr/bin/env python
"""Pin the oracle to the REAL reference (SURVEY.md 8c iii). Needs a JRE and the reference tree (class files ship in
<reference>/current); neither this image nor the GPU boxes have `java` (profiles/r02_java_probe.txt), so until someone runs
this on a machine that does, tests/test_reference_pin.py reports "parity unpinned" (xfail).

    python tools/pin_reference.py [--reference /root/reference] [--java java] [--threads 4]

For every command line of tests/pin_common.py:CASES and both main classes (jgi.BBDuk = bbdukOld.sh, bbduk.BBDukS =
bbduk.sh) it runs
    java -ea -Xmx2g -cp <reference>/current <class> in=.. [in2=..] out=.. [out2=..] outm=.. [outm2=..] stats=.. ref=adapters.fa
         ordered=t t=<threads> overwrite=t <flags>
on the seeded FASTQ files (and, for tests/pin_common.py:SEAL_CASES, `java jgi.Seal` = seal.sh on seeded pairs against a seeded
multi-sequence reference with outm / outu / stats) and stores, under tests/golden/reference_digests.json, the SHA-256 of every output file, the
`Added N kmers` line (jgi/BBDuk.java:1973), the reads/bases counters the tool prints and the java version. Also runs
resources/sample1.fq.gz + sample2.fq.gz through the cfg-2 command line (digests only; the tests regenerate nothing from
those files, they are a check for whoever has the reference tree)."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pin_common as pc  # noqa: E402


def run_case(java, cp, cls, ins, flags, workdir, threads, ref_fa):
    outs = {"out": "o1.fq", "outm": "m1.fq"}
    cmd = [java, "-ea", "-Xmx2g", "-cp", cp, cls, f"in={ins[0]}"]
    if len(ins) > 1:
        cmd.append(f"in2={ins[1]}")
        outs.update(out2="o2.fq", outm2="m2.fq")
    for key, name in outs.items():
        cmd.append(f"{key}={os.path.join(workdir, name)}")
    stats = os.path.join(workdir, "stats.txt")
    cmd += [f"stats={stats}", f"ref={ref_fa}", "ordered=t", f"t={threads}", "overwrite=t"] + flags
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return {"error": r.stderr[-2000:]}
    res = {"command": " ".join(cmd[4:])}
    for key, name in outs.items():
        path = os.path.join(workdir, name)
        res[key] = pc.sha(open(path, "rb").read()) if os.path.exists(path) else None
    m = re.search(r"Added (\d+) kmers", r.stderr)
    res["added_kmers"] = int(m.group(1)) if m else None
    for label in ("KTrimmed", "KFiltered", "KMasked", "Result"):
        m = re.search(label + r":\s+(\d+) reads \([\d.]+%\)\s+(\d+) bases", r.stderr)
        if m:
            res[label] = [int(m.group(1)), int(m.group(2))]
    if os.path.exists(stats):  # the header carries the file name and the date; keep the per-scaffold lines
        res["stats"] = pc.sha("".join(ln for ln in open(stats) if not ln.startswith("#File")).encode())
    return res


def run_seal_case(java, cp, ins, flags, workdir, threads, ref_fa):
    """java jgi.Seal (seal.sh) on the seeded pairs: digests of outm / outm2 / outu / outu2, the stats= lines, `Added N kmers`"""
    outs = {"outm": "m1.fq", "outm2": "m2.fq", "outu": "u1.fq", "outu2": "u2.fq"}
    cmd = [java, "-ea", "-Xmx2g", "-cp", cp, pc.SEAL_CLASS, f"in={ins[0]}", f"in2={ins[1]}"]
    for key, name in outs.items():
        cmd.append(f"{key}={os.path.join(workdir, name)}")
    stats = os.path.join(workdir, "stats.txt")
    cmd += [f"stats={stats}", f"ref={ref_fa}", "ordered=t", f"t={threads}", "overwrite=t"] + flags
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return {"error": r.stderr[-2000:]}
    res = {"command": " ".join(cmd[4:])}
    for key, name in outs.items():
        path = os.path.join(workdir, name)
        res[key] = pc.sha(open(path, "rb").read()) if os.path.exists(path) else None
    m = re.search(r"Added (\d+) kmers", r.stderr)
    res["added_kmers"] = int(m.group(1)) if m else None
    if os.path.exists(stats):
        res["stats"] = pc.sha("".join(ln for ln in open(stats) if not ln.startswith("#File")).encode())
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--java", default="java")
    ap.add_argument("--threads", type=int, default=4)
    a = ap.parse_args()
    if shutil.which(a.java) is None:
        sys.exit(f"no `{a.java}` on PATH: parity stays unpinned (see profiles/r02_java_probe.txt)")
    cp = os.path.join(a.reference, "current")
    if not os.path.isdir(os.path.join(cp, "jgi")):
        sys.exit(f"{cp}/jgi not found")
    ver = subprocess.run([a.java, "-version"], capture_output=True, text=True).stderr.strip().splitlines()[0]
    ref_fa = os.path.join(pc.GOLDEN, "adapters.fa")
    doc = {"java": ver, "reference": a.reference, "cases": {}}
    with tempfile.TemporaryDirectory() as tmp:
        files = {kind: pc.write_inputs(kind, tmp) for kind in ("pairs", "single", "ragged")}
        for name, (kind, flags) in pc.CASES.items():
            doc["cases"][name] = {}
            for cls in pc.MAIN_CLASSES:
                wd = os.path.join(tmp, name + "_" + cls)
                os.makedirs(wd)
                doc["cases"][name][cls] = run_case(a.java, cp, cls, files[kind], flags, wd, a.threads, ref_fa)
                print(name, cls, doc["cases"][name][cls].get("added_kmers"), file=sys.stderr)
        s1, s2 = (os.path.join(a.reference, "resources", f) for f in ("sample1.fq.gz", "sample2.fq.gz"))
        if os.path.exists(s1) and os.path.exists(s2):
            doc["samples"] = {}
            for cls in pc.MAIN_CLASSES:
                wd = os.path.join(tmp, "samples_" + cls)
                os.makedirs(wd)
                doc["samples"][cls] = run_case(a.java, cp, cls, [s1, s2], pc.CASES["cfg2_ktrim_r_k23_mink11_hdist1_tpe"][1], wd,
                                               a.threads, os.path.join(a.reference, "resources", "adapters.fa"))
        doc["seal_cases"] = {}
        seal_ref, seal_ins = pc.write_seal_inputs(tmp)
        for name, flags in pc.SEAL_CASES.items():
            wd = os.path.join(tmp, name)
            os.makedirs(wd)
            doc["seal_cases"][name] = run_seal_case(a.java, cp, seal_ins, flags, wd, a.threads, seal_ref)
            print(name, doc["seal_cases"][name].get("added_kmers"), file=sys.stderr)
    with open(pc.DIGESTS, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    print("wrote", pc.DIGESTS)


if __name__ == "__main__":
    main()
