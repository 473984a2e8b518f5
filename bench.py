#!/usr/bin/env python
"""bench.py -- BBDuk reads/s on synthetic 2x150 bp pairs, k=23 mink=11 hdist=1 ktrim=r tpe (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            the CPU restatement of the reference on host cores

A step = one pass of the hot path over one batch of read pairs per GPU. `value` is whole-job reads/s with
the batches already resident in HBM (device generator, no PCIe); `e2e` is the same metric through the
C-ABI call with pinned HOST buffers (H2D + kernels + D2H inside the timed region). The k-mer table is
built on rank 0 and replicated with one NCCL broadcast per blob; reads shard across ranks with no
per-read collective (scaling = weak). Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

READ_LEN = 150
CFG = dict(k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1)
WORKLOAD = "bbduk.sh ktrim=r k=23 mink=11 hdist=1 tpe, ref=adapters.fa, synthetic 2x150 bp PE (cfg 2; tbo stays on the host)"
ALG_BYTES_PER_READ = READ_LEN + 4 + 8  # bases + 4 B offset in + 8 B result out (hi + id0); table on-chip
FALLBACK_HBM_GBS = 6650.0


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def adapters_ref():
    from bbtools_b200.fasta import read_fasta
    _, b, off = read_fasta(os.path.join(ROOT, "tests", "golden", "adapters.fa"))
    return b, off


def cpu_reference_rate(n_pairs, threads, seed=1, first_pair=0, repeats=1):
    """the oracle (CPU restatement of the reference's Java loop) on a bounded sample -> reads/s"""
    from bbtools_b200 import make_cfg, synth
    from oracle.oracle import Oracle
    rb, roff = adapters_ref()
    o = Oracle(make_cfg(**CFG))
    o.add_ref(rb, roff)
    o.finalize()
    bases, offsets = synth.paired_adapter_reads(n_pairs, first_pair=first_pair, read_len=READ_LEN, seed=seed)
    o.process(bases[: 2 * READ_LEN * 2000], offsets[:4001], True, threads=threads)  # warm the table into cache
    best = 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        o.process(bases, offsets, True, threads=threads)
        dt = time.perf_counter() - t0
        best = max(best, 2 * n_pairs / dt)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_pairs = args.ref_pairs
    times = []
    from bbtools_b200 import make_cfg, synth
    from oracle.oracle import Oracle
    rb, roff = adapters_ref()
    o = Oracle(make_cfg(**CFG))
    o.add_ref(rb, roff)
    o.finalize()
    bases, offsets = synth.paired_adapter_reads(n_pairs, read_len=READ_LEN, seed=1)
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o.process(bases, offsets, True, threads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = 2 * n_pairs * len(times) / total
    sample = f"{2 * n_pairs} reads ({n_pairs} synthetic 2x150 pairs, seed 1) per step, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "bbduk_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": n_pairs,
                   "note": "CPU restatement (C port) of jgi.BBDuk's k-mer block on host cores; no JVM in the image"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from bbtools_b200 import _lib, make_cfg
    from bbtools_b200.bbduk import BBDukIndexGPU

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the BBDuk path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # ---- table: built on rank 0, replicated by NCCL broadcast (no per-read collective afterwards) ----
    cfg = make_cfg(device=local_rank, **CFG)
    eng = BBDukIndexGPU(cfg)
    t_build = time.perf_counter()
    if rank == 0:
        rb, roff = adapters_ref()
        eng.add_ref(rb, roff)
        stored = eng.finalize()
    if world > 1:
        stored = eng.broadcast_table(src=0)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    assert stored == 217135, stored
    eng.set_max_read_len(READ_LEN)

    # ---- HBM-resident batches from the device generator (distinct per rank and per buffer) ----------
    n_pairs = args.pairs_per_step
    n_reads = 2 * n_pairs
    nbuf = 2
    bufs = []
    for b in range(nbuf):
        d_bases = torch.empty(n_reads * READ_LEN, dtype=torch.uint8, device=dev)
        d_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
        first = (rank * nbuf + b) * n_pairs
        rc = lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), n_pairs, first, READ_LEN, C.c_uint64(1), 50, 5, None)
        assert rc == 0
        bufs.append((d_bases, d_off))
    outs = {"id0": torch.empty(n_reads, dtype=torch.int32, device=dev),
            "hi": torch.empty(n_reads, dtype=torch.int32, device=dev),
            "flags": torch.empty(n_reads, dtype=torch.uint8, device=dev)}
    d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)

    def step(i):
        d_bases, d_off = bufs[i % nbuf]
        eng.process_device(d_bases, d_off, n_reads, True, outs, d_stats=d_stats, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
            ev[i + 1].record(stream)
    barrier()
    launches = eng.launches - l0
    clocks = sampler.stop()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = world * n_reads * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers ------------------------------------
    e_pairs = args.e2e_pairs
    e_reads = 2 * e_pairs
    h_bases = torch.empty(e_reads * READ_LEN, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(bufs[0][0][: e_reads * READ_LEN])
    h_off = torch.arange(0, (e_reads + 1) * READ_LEN, READ_LEN, dtype=torch.int64).pin_memory()
    from bbtools_b200._abi import Outputs
    hout = Outputs(0)
    hout.n = e_reads
    pinned = {"id0": torch.empty(e_reads, dtype=torch.int32, pin_memory=True),
              "hi": torch.empty(e_reads, dtype=torch.int32, pin_memory=True),
              "flags": torch.empty(e_reads, dtype=torch.uint8, pin_memory=True)}
    hout.id0, hout.hi, hout.flags = (pinned[k].numpy() for k in ("id0", "hi", "flags"))
    hout.id0b = hout.lo = hout.count = None
    hb, ho = h_bases.numpy(), h_off.numpy()
    h2d = int(hb.nbytes + ho.nbytes)
    d2h = int(sum(pinned[k].numpy().nbytes for k in pinned))
    e_steps = max(2, min(args.steps, 8))
    for _ in range(2):
        eng.process(hb, ho, True, out=hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        _, est = eng.process(hb, ho, True, out=hout)
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_value = world * e_reads * e_steps / float(te.item())

    # ---- sanity: the timed batch agrees with the CPU oracle on a slice (checker only) -----------------
    parity = None
    if rank == 0:
        from bbtools_b200 import synth
        from oracle.oracle import Oracle
        chk = 20000
        o = Oracle(make_cfg(**CFG))
        rb, roff = adapters_ref()
        o.add_ref(rb, roff)
        o.finalize()
        hb2, ho2 = synth.paired_adapter_reads(chk, first_pair=(args.warmup + args.steps - 1) % nbuf * n_pairs, seed=1)
        want, _ = o.process(hb2, ho2, True, threads=min(8, os.cpu_count() or 1))
        step(args.warmup + args.steps - 1)
        torch.cuda.synchronize()
        parity = bool(np.array_equal(outs["hi"][: 2 * chk].cpu().numpy(), want.hi) and
                      np.array_equal(outs["id0"][: 2 * chk].cpu().numpy(), want.id0) and
                      np.array_equal(outs["flags"][: 2 * chk].cpu().numpy(), want.flags))

    if rank == 0:
        peak, peak_kind = load_peak()
        kern_ms = statistics.mean(step_ms)
        achieved = n_reads * ALG_BYTES_PER_READ / (kern_ms * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        cpu_pairs = args.cpu_pairs
        cpu_rate = None
        if world == 1 and cpu_pairs > 0:
            # the sample is the head of the timed workload, copied back from the device generator
            from bbtools_b200 import make_cfg as _mk
            from oracle.oracle import Oracle as _Or
            cpu_pairs = min(cpu_pairs, n_pairs)
            cb = bufs[0][0][: 2 * cpu_pairs * READ_LEN].cpu().numpy()
            co = np.arange(0, (2 * cpu_pairs + 1) * READ_LEN, READ_LEN, dtype=np.int64)
            oo = _Or(_mk(**CFG))
            rb, roff = adapters_ref()
            oo.add_ref(rb, roff)
            oo.finalize()
            oo.process(cb[: 4000 * READ_LEN], co[:4001], True, threads=cores)
            t0 = time.perf_counter()
            oo.process(cb, co, True, threads=cores)
            cpu_rate = 2 * cpu_pairs / (time.perf_counter() - t0)
        line = {
            "metric": "bbduk_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": n_pairs, "read_len": READ_LEN,
                       "stored_kmers": stored, "l2": f"inputs {n_reads * READ_LEN / 2**20:.0f} MiB per step > 126 MB L2, "
                       f"{nbuf} alternating buffers, no flush", "table_build_s": round(t_build, 3),
                       "parity_vs_oracle_on_timed_batch": parity},
            "e2e": {"value": e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pairs_per_step_per_gpu": e_pairs, "steps": e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_kind, "kernel": "bbduk_fast_kernel",
                         "algorithmic_bytes_per_read": ALG_BYTES_PER_READ, "ms_per_launch": kern_ms},
        }
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": "reads/s", "cores": cores, "kind": "port",
                                    "sample": f"{2 * cpu_pairs} reads of the same synthetic workload (seed 1), oracle C port, "
                                              f"{cores} threads"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=8 << 20, help="pairs per GPU per step (HBM-resident)")
    ap.add_argument("--e2e-pairs", type=int, default=2 << 20, help="pairs per GPU per end-to-end step (host buffers)")
    ap.add_argument("--cpu-pairs", type=int, default=4 << 20, help="pairs of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--ref-pairs", type=int, default=1 << 19, help="pairs per step of --impl reference")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
