"""Times seal_match_kernel on a synthetic workload (device buffers, CUDA events) and checks a slice against the oracle.

  python tools/time_seal.py [--refs 2000] [--ref-len 5000] [--pairs 1048576] [--iters 5] [--check 20000]

Reads are 150 bp pieces of the references (either strand, 1 % substitutions, 0.1 % N). --mode strains: groups of four
references 2 % apart (k-mers with 1-4 ids); --mode many: a quarter of the references are pieces of one sequence (value
lists of hundreds of ids, the worst case for the per-pair lists). Prints one JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bbtools_b200 import _lib  # noqa: E402
from bbtools_b200 import seal as PS  # noqa: E402


def workload(n_refs, ref_len, n_pairs, seed=1, mode="strains", read_seed=None):
    """-> (refs [n_refs, ref_len] uint8, reads [2 * n_pairs, 150] uint8); read_seed: another read sample of the same references"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    refs = np.empty((n_refs, ref_len), np.uint8)
    if mode == "strains":  # groups of four references 2 % apart: k-mers carry 1-4 ids
        for r in range(n_refs):
            if r % 4 == 0:
                refs[r] = acgt[rng.integers(0, 4, ref_len)]
            else:
                refs[r] = refs[r - r % 4]
                q = rng.integers(0, ref_len, max(1, ref_len // 50))
                refs[r, q] = acgt[rng.integers(0, 4, len(q))]
    else:  # "many": every fourth reference is a 1 %-mutated piece of ONE sequence: value lists of hundreds of ids
        common = acgt[rng.integers(0, 4, ref_len + ref_len // 2)]
        for r in range(n_refs):
            if r % 4 == 0:
                p = int(rng.integers(0, ref_len // 2))
                refs[r] = common[p:p + ref_len]
                q = rng.integers(0, ref_len, max(1, ref_len // 100))
                refs[r, q] = acgt[rng.integers(0, 4, len(q))]
            else:
                refs[r] = acgt[rng.integers(0, 4, ref_len)]
    if read_seed is not None:
        rng = np.random.default_rng(read_seed)
    n = 2 * n_pairs
    src = rng.integers(0, n_refs, n_pairs).repeat(2)
    pos = rng.integers(0, ref_len - 150, n)
    idx = pos[:, None] + np.arange(150)[None, :]
    mat = refs[src[:, None], idx]
    err = rng.random((n, 150))
    mat = np.where(err < 0.01, acgt[rng.integers(0, 4, (n, 150))], mat)
    mat = np.where(err > 0.999, np.uint8(ord("N")), mat)
    comp = np.zeros(256, np.uint8)
    for a, c in zip(b"ACGTN", b"TGCAN"):
        comp[a] = c
    flip = rng.integers(0, 2, n).astype(bool)
    mat[flip] = comp[mat[flip]][:, ::-1]
    return refs, np.ascontiguousarray(mat)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refs", type=int, default=2000)
    ap.add_argument("--ref-len", type=int, default=5000)
    ap.add_argument("--pairs", type=int, default=1 << 20)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--check", type=int, default=20000)
    ap.add_argument("--mode", default="strains", choices=["strains", "many"])
    a = ap.parse_args()
    refs, mat = workload(a.refs, a.ref_len, a.pairs, mode=a.mode)
    n = mat.shape[0]
    cfg = PS.make_cfg(ambig_mode=PS.AMBIG_RANDOM)
    g = PS.SealIndexGPU(cfg)
    roff = np.arange(a.refs + 1, dtype=np.int64) * a.ref_len
    g.add_ref(refs.reshape(-1), roff)
    t0 = time.time()
    stored, entries, refk = g.finalize()
    t_build = time.time() - t0
    lib = _lib.load()
    dev = torch.device("cuda:0")
    d_b = torch.from_numpy(np.concatenate([mat.reshape(-1), np.zeros(16, np.uint8)])).to(dev)
    off = np.arange(n + 1, dtype=np.int64) * 150
    d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
    nu = n // 2
    stride = cfg.ids_stride
    d_res = torch.zeros(nu * (4 + stride), dtype=torch.int32, device=dev)
    d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
    out = PS.SealOut()
    base = d_res.data_ptr()
    out.n_assigned, out.first_id, out.n_sites, out.max_hits, out.ids = base, base + 4 * nu, base + 8 * nu, base + 12 * nu, base + 16 * nu
    st = torch.cuda.current_stream().cuda_stream

    def run():
        rc = lib.seal_b200_process_device(g.h, d_b.data_ptr(), d_off.data_ptr(), n, 1, 0, C.byref(out), d_stats.data_ptr(), st)
        assert rc == 0, lib.seal_b200_last_error(g.h)

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    r = d_res.cpu().numpy()
    line = {"workload": f"seal k=31 mm=t ambig=random, {a.refs} refs x {a.ref_len} bp ({a.mode}), {a.pairs} pairs x 2 x 150 bp",
            "stored_kmers": stored, "entries": entries, "build_s": round(t_build, 3), "ms_per_batch": round(ms, 3),
            "reads_per_s": round(n / (ms * 1e-3)), "assigned_pairs": int((r[:nu] > 0).sum()), "ambiguous_pairs": int((r[2 * nu:3 * nu] > 1).sum())}
    # algorithmic bytes per read: its bases + one 32-byte bucket sector per probed k-mer (L - k + 1) + 12 B of results
    per_read = 150 + (150 - cfg.k + 1) * 32 + 12
    try:
        peak, src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        peak, src = 7700.0, "fallback"
    ach = per_read * n / (ms * 1e-3) / 1e9
    line["roofline"] = {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "peak_source": src, "unit": "GB/s",
                        "frac": round(ach / peak, 4), "algorithmic_bytes_per_read": per_read}
    if a.check > 0:
        from oracle import seal as S
        o = S.SealOracle(cfg)
        o.add_ref(refs.reshape(-1), roff)
        o.finalize()
        m = min(a.check, nu) * 2
        t0 = time.time()
        want, _ = o.process(mat[:m].reshape(-1), off[:m + 1], True, 0)
        line["oracle_reads_per_s_1core"] = round(m / (time.time() - t0))
        h = m // 2
        ok = (np.array_equal(r[:h], want.n_assigned) and np.array_equal(r[nu:nu + h], want.first_id)
              and np.array_equal(r[2 * nu:2 * nu + h], want.n_sites) and np.array_equal(r[3 * nu:3 * nu + h], want.max_hits)
              and np.array_equal(r[4 * nu:4 * nu + h * stride], want.ids))
        line["parity_vs_oracle_pairs"] = h
        line["parity"] = bool(ok)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
