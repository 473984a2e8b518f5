#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep (compiled with -lineinfo):
    python tools/ncu_lines.py report.ncu-rep [n_tiles] [top]
Prints warp instructions executed per source line (and per tile if n_tiles is given), largest first."""
import csv
import subprocess
import sys

rep = sys.argv[1]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None
agg = {}
hdr = None
for row in csv.reader(txt.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if hdr and row[0].strip().isdigit():
        try:
            inst = float(row[i_inst])
            samp = float(row[i_samp])
        except ValueError:
            continue
        key = (cur_file, int(row[0]))
        a = agg.setdefault(key, [0.0, 0.0, row[1].strip()[:110]])
        a[0] += inst
        a[1] += samp
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot:.0f}" + (f" = {tot / tiles:.0f} per tile" if tiles else "") + f", samples {tots:.0f}")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    per = f"{v[0] / tiles:8.1f}" if tiles else f"{v[0]:12.0f}"
    print(f"{per} {100 * v[0] / tot:5.1f}% samp {100 * v[1] / max(tots, 1):5.1f}%  {f}:{ln}  {v[2]}")

# optional phase summary for probe_fast2.cu (line ranges of the phases; pass --phases as last argument)
if "--phases" in sys.argv:
    import re
    src = open(__file__.rsplit("/tools/", 1)[0] + "/bbtools_b200/csrc/probe_fast2.cu").read().splitlines()
    marks = []
    for i, ln in enumerate(src, 1):
        m = re.search(r"// ---- ([A-Z][0-9]?)\. ", ln)
        if m:
            marks.append((i, m.group(1)))
        if "auto drain = " in ln:
            marks.append((i, "C-drain"))
        if "bool done = !scan" in ln:
            marks.append((i, "C-rounds"))
        if "auto cand_word" in ln or "auto und_word" in ln:
            marks.append((i, "C-candword"))
        if "int found = 0, id0" in ln:
            marks.append((i, "C-ktriml"))
    marks.sort()
    ph = {}
    phs = {}
    for (f, ln), v in agg.items():
        if f == "probe_fast2.cu":
            name = "setup"
            for (l0, nm) in marks:
                if ln >= l0:
                    name = nm
        elif f == "fast_common.cuh":
            name = "classify" if ln < 34 else "stream" if ln < 53 else "exact/spread"
        elif f == "bbduk_dev.cuh":
            name = "table_get/rcomp"
        else:
            name = f
        ph[name] = ph.get(name, 0) + v[0]
        phs[name] = phs.get(name, 0) + v[1]
    print("--- phases")
    for k_, v in sorted(ph.items(), key=lambda kv: -kv[1]):
        print(f"{v / tiles:8.1f} {100 * v / tot:5.1f}%  samples {100 * phs[k_] / max(tots, 1):5.1f}%  {k_}")
