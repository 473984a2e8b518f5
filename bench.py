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
WORKLOAD = ("bbduk.sh ktrim=r k=23 mink=11 hdist=1 tpe, ref=adapters.fa, synthetic 2x150 bp PE (cfg 2: the k-mer block; "
            "config.kmer_block_plus_tbo times the same step followed by the tbo kernel)")
ALG_BYTES_PER_READ = READ_LEN + 4 + 8  # bases + 4 B offset in + 8 B result out (hi + id0); table on-chip
FALLBACK_HBM_GBS = 6650.0
# dram__bytes_read.sum + dram__bytes_write.sum per read of the dominant kernel, from the committed `ncu --set full`
# captures of the same workloads (a profiler cannot run inside the timed bench; the captures are re-taken whenever the
# kernel changes and the JSON line names the file: roofline.traffic_source);
# cfg 5: profiles/r01_e_kcount_kernel.txt: 28.33 GB + 6.91 GB for 217.6 M k-mers = 162 B per k-mer
NCU_TRAFFIC_BYTES_PER_READ = {"cfg2": (1299953000 + 75363840) / 8388608,  # profiles/r02n_fast2_kernel_raw.txt: an 8,388,608-read launch of bbduk_fast2_kernel
                              # profiles/r02_cfg3_direct_kernel_details.txt / r02_cfg4_...: 4,194,304-read launches of bbduk_direct_kernel
                              "cfg3": (30556543000 + 93049088) / 4194304, "cfg4": (35478975000 + 113840128) / 4194304,
                              "cfg5": 120 * (28328275000 + 6911184000) / 217637790}


TRAFFIC_SOURCE = {"cfg2": "profiles/r02n_fast2_kernel_raw.txt", "cfg3": "profiles/r02_cfg3_direct_kernel_details.txt",
                  "cfg4": "profiles/r02_cfg4_direct_kernel_details.txt", "cfg5": "profiles/r01_e_kcount_kernel.txt"}


# Libraries write to the process's stdout behind Python's back (NCCL prints its version line there when a communicator
# is created), and the contract is ONE JSON line on stdout: file descriptor 1 is pointed at stderr for the whole run and
# the line goes to a private duplicate of the original stdout.
_JSON_OUT = None


def claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


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
                                          "--format=csv,noheader,nounits", "-lms", "50"],
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


def pin_to_gpu_numa_node(local_rank, ranks_on_node):
    """One process per GPU: keep this rank's threads (and so its first-touch host buffers and the library's packing pool, whose
    threads inherit the mask) on the NUMA node the GPU hangs off, instead of letting 8 ranks roam over both sockets.
    BBDUK_B200_NUMA=0 switches it off. Returns a short description for the JSON line."""
    if os.environ.get("BBDUK_B200_NUMA", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return "off"
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "no numa node reported"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if len(allowed) < 2:
            return f"node {node}: too few cpus"
        os.sched_setaffinity(0, allowed)
        return f"node {node}: {len(allowed)} cpus"
    except Exception as e:  # a VM without the sysfs entries: leave the scheduler alone
        return f"unavailable ({type(e).__name__})"


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


def seal_inputs(wl, n_pairs, read_seed=None):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import time_seal
    n_refs, ref_len, seed = wl["refs"]
    refs, mat = time_seal.workload(n_refs, ref_len, n_pairs, seed=seed, read_seed=read_seed)
    return refs.reshape(-1), np.arange(n_refs + 1, dtype=np.int64) * ref_len, mat


def run_reference_seal(args):
    """--impl reference --workload seal: the Seal oracle (C restatement of jgi.Seal's matching block, one ProcessThread per
    host core) on a bounded sample of the same workload"""
    from bbtools_b200 import seal as PS
    from oracle import seal as S
    wl = WORKLOADS["seal"]
    cores = os.cpu_count() or 1
    n_pairs = min(args.ref_pairs, 20000 * cores)
    rb, roff, mat = seal_inputs(wl, n_pairs)
    o = S.SealOracle(PS.make_cfg())
    o.add_ref(rb, roff)
    o.finalize()
    off = np.arange(2 * n_pairs + 1, dtype=np.int64) * READ_LEN
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o.process(mat.reshape(-1), off, True, 0, threads=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = 2 * n_pairs * len(times) / total
    sample = f"{2 * n_pairs} reads ({n_pairs} pairs of the same synthetic workload) per step, {cores} threads"
    emit(json.dumps({
        "impl": "reference", "metric": "seal_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": wl["desc"], "pairs_per_step": n_pairs,
                   "note": "CPU restatement (C port, sorted entries behind a hash index as the map) of jgi.Seal's matching block; no JVM in the image"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "seal":
        return run_reference_seal(args)
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
    emit(json.dumps({
        "impl": "reference", "metric": "bbduk_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": n_pairs,
                   "note": "CPU restatement (C port) of jgi.BBDuk's k-mer block on host cores; no JVM in the image"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- workloads (BASELINE.json configs) ---------------------------------------------------
# alg_bytes: ALGORITHMIC bytes per read = L + 12 + 8*P*[table in HBM]
WORKLOADS = {
    "cfg2": dict(cfg=CFG, paired=True, read_len=READ_LEN, alg_bytes=READ_LEN + 4 + 8, kernel="bbduk_fast2_kernel",
                 desc=WORKLOAD, stored=217135),
    "cfg3": dict(cfg=dict(k=31), paired=False, read_len=READ_LEN, alg_bytes=READ_LEN + 12 + 8 * 120, kernel="bbduk_direct_kernel",
                 desc="bbduk.sh k=31 (kfilter, mm=t) vs 100 x 1 Mbp synthetic reference (seed 7), synthetic 150 bp SE reads, "
                      "10 % contaminant with 1 % subs (cfg 3)", ref=(100, 1_000_000, 7)),
    "cfg4": dict(cfg=dict(k=27, hdist=2), paired=True, read_len=READ_LEN, alg_bytes=READ_LEN + 12 + 8 * 124,
                 kernel="bbduk_direct_kernel",
                 desc="bbduk.sh k=27 hdist=2 (kfilter) vs 100 x 1 kbp synthetic reference (seed 9), synthetic 2x150 bp PE as cfg 2 (cfg 4)",
                 ref=(100, 1000, 9)),
    "cfg5": dict(paired=False, read_len=READ_LEN, alg_bytes=READ_LEN + 4 + 8 * 120, kernel="kcount_kernel",
                 desc="kmercountexact.sh k=31, synthetic 150 bp SE reads sampled from a 100 Mbp genome (seed 11), 0.1 % subs (cfg 5)",
                 genome=100_000_000),
    # SURVEY.md 8f row 4 (no BASELINE.json config): Seal's multi-value table + per-pair assignment, tools/time_seal.py's workload
    "seal": dict(paired=True, read_len=READ_LEN, alg_bytes=READ_LEN + 12 + 8 * 120, kernel="seal_match_kernel",
                 desc="seal.sh k=31 (mm=t, ambig=random, kpt=t) vs 2000 x 5 kbp synthetic references in strain groups of four 2 % apart "
                      "(seed 1), synthetic 2x150 bp PE sampled from them, 1 % subs, 0.1 % N (SURVEY 8f row 4)",
                 refs=(2000, 5000, 1)),
}


def table_checksum(eng, torch):
    """order-independent checksum of the (key, id) pairs of the device hash array + the filter images, computed on the
    device in slices -> (63-bit checksum, number of keys). Equal on all ranks <=> the replicas hold the same table."""
    d = eng.table_describe()
    keys = eng._view(d.d_keys, d.n_slots * 8).view(torch.int64)
    vals = eng._view(d.d_vals, d.n_slots * 4).view(torch.int32)
    filt = eng._view(d.d_filter, d.n_filter_words * 4).view(torch.int32)
    acc, n_keys = 0, 0
    step = 1 << 26
    for a in range(0, d.n_slots, step):
        k = keys[a:a + step]
        m = k != -1
        x = k * -7046029254386353131 + vals[a:a + step].to(torch.int64) * -4417276706812531889
        acc = (acc + int(torch.where(m, x, torch.zeros_like(x)).sum().item())) & 0xFFFFFFFFFFFFFFFF
        n_keys += int(m.sum().item())
    for a in range(0, d.n_filter_words, step):
        f = filt[a:a + step].to(torch.int64)
        idx = torch.arange(a, a + f.numel(), device=f.device, dtype=torch.int64)
        acc = (acc + int((f * (2 * idx + 1)).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return acc >> 1, n_keys


def build_engine(wl, rank, local_rank, world, lib, torch):
    """table built on rank 0 and replicated by NCCL broadcast; returns (engine, stored, d_ref or None)"""
    from bbtools_b200 import make_cfg
    from bbtools_b200.bbduk import BBDukIndexGPU
    cfg = make_cfg(device=local_rank, **wl["cfg"])
    eng = BBDukIndexGPU(cfg)
    d_ref = None
    if "ref" in wl:
        n_scaf, scaf_len, seed = wl["ref"]
        d_ref = torch.empty(n_scaf * scaf_len, dtype=torch.uint8, device="cuda")
        assert lib.bbduk_b200_synth_reference(d_ref.data_ptr(), d_ref.numel(), C.c_uint64(seed), None) == 0
        torch.cuda.synchronize()
    stored = 0
    if rank == 0:
        if d_ref is None:
            rb, roff = adapters_ref()
        else:
            rb = d_ref.cpu().numpy()
            roff = np.arange(0, rb.size + 1, wl["ref"][1], dtype=np.int64)
        eng.add_ref(rb, roff)
        stored = eng.finalize()
    if world > 1:
        stored = eng.broadcast_table(src=0)
    torch.cuda.synchronize()
    return eng, stored, d_ref


def run_kcount(args, wl):
    """config 5: a step = counting one batch of reads into the rank-private table; at N>1 the ONE exchange
    (all_to_all by key owner) runs once after the timed steps and is reported separately"""
    import torch
    import torch.distributed as dist

    from bbtools_b200 import _lib
    from bbtools_b200.kcount import KmerTableSetGPU, exchange_counts, global_summary
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    L, G = wl["read_len"], wl["genome"]
    n_reads = 2 * args.pairs_per_step
    tab = KmerTableSetGPU(31, True, initial_keys=1 << 29, device=local_rank)
    nbuf = 2
    bufs = []
    for b in range(nbuf):
        d_bases = torch.empty(n_reads * L, dtype=torch.uint8, device=dev)
        d_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
        first = (rank * nbuf + b) * n_reads
        assert lib.kcount_b200_synth_reads(d_bases.data_ptr(), d_off.data_ptr(), n_reads, first, L, G, C.c_uint64(11), 10, None) == 0
        bufs.append((d_bases, d_off))
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)

    def step(i):
        d_bases, d_off = bufs[i % nbuf]
        tab.add_reads_device(d_bases, d_off, n_reads, n_reads * L, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = tab.table_info()["launches"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
            ev[i + 1].record(stream)
    barrier()
    launches = tab.table_info()["launches"] - l0
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = world * n_reads * args.steps / (total_ms_max * 1e-3)
    # end to end: host buffers through kcount_b200_add_reads
    e_reads = 2 * args.e2e_pairs
    h_bases = torch.empty(e_reads * L, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(bufs[0][0][: e_reads * L])
    hb = h_bases.numpy()
    ho = np.arange(0, (e_reads + 1) * L, L, dtype=np.int64)
    tab.add_reads(hb, ho)
    barrier()
    t0 = time.perf_counter()
    e_steps = 3
    for _ in range(e_steps):
        tab.add_reads(hb, ho)
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_value = world * e_reads * e_steps / float(te.item())
    exch_ms = None
    st = tab.stats()
    if world > 1:
        barrier()
        t0 = time.perf_counter()
        owner = exchange_counts(tab)
        uniq, _ = global_summary(owner, 1000)
        barrier()
        exch_ms = 1e3 * (time.perf_counter() - t0)
    else:
        uniq = st["unique_kmers"]
    cpu_line = None
    if rank == 0 and world == 1 and args.cpu_pairs > 0:
        # the counting oracle (CPU restatement of KmerTableSet's loop, single-threaded by construction) on the head of the batch
        from oracle.kcount import KCountOracle
        c_reads = min(n_reads, 1 << 20)
        cb = bufs[0][0][: c_reads * L].cpu().numpy()
        co = np.arange(0, (c_reads + 1) * L, L, dtype=np.int64)
        ko = KCountOracle(31, True)
        t0 = time.perf_counter()
        ko.add_reads(cb, co)
        cpu_line = {"value": c_reads / (time.perf_counter() - t0), "unit": "reads/s", "cores": 1, "kind": "port",
                    "sample": f"{c_reads} reads of the same synthetic workload, counting oracle (C port), 1 thread"}
        del ko
    if rank == 0:
        peak, peak_kind = load_peak()
        kern_ms = total_ms_max / args.steps
        achieved = n_reads * wl["alg_bytes"] / (kern_ms * 1e-3) / 1e9
        info = tab.table_info()
        emit(json.dumps({
            **({"cpu_baseline": cpu_line} if cpu_line else {}),
            "metric": "kmercount_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": wl["desc"], "reads_per_step_per_gpu": n_reads, "read_len": L, "unique_kmers": uniq,
                       "kmers_per_s": value * (L - 30), "table_bytes_per_gpu": info["bytes"], "exchange_ms": exch_ms,
                       "l2": f"table {info['bytes'] / 2**30:.0f} GiB >> 126 MB L2; inputs {n_reads * L / 2**20:.0f} MiB per step"},
            "e2e": {"value": e_value, "unit": "reads/s", "h2d_bytes_per_step": int(hb.nbytes + (e_reads + 1) * 4),
                    "d2h_bytes_per_step": 32, "reads_per_step_per_gpu": e_reads, "steps": e_steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_READ["cfg5"] * n_reads,
                         "traffic_unit": "bytes per launch (ncu dram read+write per k-mer x k-mers per launch)",
                         "peak_source": peak_kind, "kernel": wl["kernel"],
                         "algorithmic_bytes_per_read": wl["alg_bytes"], "ms_per_launch": kern_ms,
                         "dram_line_bytes_per_read": L + 4 + (128 + 32) * 120},
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_seal(args, wl):
    """a step = seal_match_kernel over one HBM-resident batch of pairs (every rank its own read sample,
    its own replica of the table built from the same references; no per-read collective)"""
    import torch
    import torch.distributed as dist

    from bbtools_b200 import _lib
    from bbtools_b200 import seal as PS
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    L = wl["read_len"]
    n_pairs = min(args.pairs_per_step, 1 << 20)
    n_reads = 2 * n_pairs
    rb, roff, mat = seal_inputs(wl, n_pairs, read_seed=None if rank == 0 else 100 + rank)
    cfg = PS.make_cfg(device=local_rank)
    eng = PS.SealIndexGPU(cfg)
    eng.add_ref(rb, roff)
    t0 = time.perf_counter()
    stored, entries, _ = eng.finalize()
    build_s = time.perf_counter() - t0
    off = np.arange(n_reads + 1, dtype=np.int64) * L
    d_b = torch.from_numpy(np.concatenate([mat.reshape(-1), np.zeros(16, np.uint8)])).to(dev)
    d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
    nu, stride = n_pairs, cfg.ids_stride
    d_res = torch.zeros(nu * (4 + stride), dtype=torch.int32, device=dev)
    d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
    out = PS.SealOut()
    base = d_res.data_ptr()
    out.n_assigned, out.first_id, out.n_sites, out.max_hits, out.ids = base, base + 4 * nu, base + 8 * nu, base + 12 * nu, base + 16 * nu
    stream = torch.cuda.Stream(device=dev)

    def step(i):
        rc = lib.seal_b200_process_device(eng.h, d_b.data_ptr(), d_off.data_ptr(), n_reads, 1, 0, C.byref(out), d_stats.data_ptr(),
                                          stream.cuda_stream)
        if rc:
            raise RuntimeError(lib.seal_b200_last_error(eng.h).decode())

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
    total_ms = ev[0].elapsed_time(ev[args.steps])
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = world * n_reads * args.steps / (total_ms_max * 1e-3)
    # every rank checks the head of ITS timed batch against the oracle; AND over ranks
    from oracle import seal as S
    ora = S.SealOracle(cfg)
    ora.add_ref(rb, roff)
    ora.finalize()
    cores = os.cpu_count() or 1
    h = min(4000 * cores, n_pairs)
    t0 = time.perf_counter()
    want, _ = ora.process(mat[:2 * h].reshape(-1), off[:2 * h + 1], True, 0, threads=cores)
    cpu_dt = time.perf_counter() - t0
    r = d_res.cpu().numpy()
    ok = (np.array_equal(r[:h], want.n_assigned) and np.array_equal(r[nu:nu + h], want.first_id)
          and np.array_equal(r[2 * nu:2 * nu + h], want.n_sites) and np.array_equal(r[3 * nu:3 * nu + h], want.max_hits)
          and np.array_equal(r[4 * nu:4 * nu + h * stride], want.ids))
    okt = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    # end to end: host buffers through seal_b200_process (H2D + kernel + D2H inside the timed region)
    e_pairs = min(args.e2e_pairs, n_pairs, 1 << 19)
    hb, ho = mat[:2 * e_pairs].reshape(-1), off[:2 * e_pairs + 1]
    eng.process(hb, ho, True, 0)
    barrier()
    e_steps = 3
    t0 = time.perf_counter()
    for _ in range(e_steps):
        eng.process(hb, ho, True, 0)
    e_dt = time.perf_counter() - t0
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_value = world * 2 * e_pairs * e_steps / float(te.item())
    if rank == 0:
        peak, peak_kind = load_peak()
        kern_ms = total_ms_max / args.steps
        achieved = n_reads * wl["alg_bytes"] / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "seal_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": wl["desc"], "pairs_per_step_per_gpu": n_pairs, "stored_kmers": stored, "table_entries": entries,
                       "table_build_s": build_s, "parity_vs_oracle_on_timed_batch": bool(int(okt.item())), "parity_pairs_per_rank": h,
                       "l2": f"inputs {n_reads * L / 2**20:.0f} MiB per step > 126 MB L2; table {stored * 2 * 16 / 2**20:.0f} MiB"},
            "e2e": {"value": e_value, "unit": "reads/s", "h2d_bytes_per_step": int(hb.nbytes + (2 * e_pairs + 1) * 4),
                    "d2h_bytes_per_step": int(e_pairs * (4 + stride) * 4 + 64), "pairs_per_step_per_gpu": e_pairs, "steps": e_steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (13122799000 + 48030976) / 524288 * n_reads,
                         "traffic_unit": "bytes per launch (ncu dram read+write per read x reads per launch)",
                         "traffic_source": "profiles/r02w_seal_match_kernel_raw.txt (the kernel before keys and values shared a line)",
                         "peak_source": peak_kind, "kernel": wl["kernel"], "algorithmic_bytes_per_read": wl["alg_bytes"],
                         "ms_per_launch": kern_ms},
        }
        if world == 1:
            line["cpu_baseline"] = {"value": 2 * h / cpu_dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                    "sample": f"{2 * h} reads of the timed batch, Seal oracle (C port), {cores} threads"}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist

    from bbtools_b200 import _lib, make_cfg

    wl = WORKLOADS[args.workload]
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the BBDuk path has no CPU fallback")
    if args.workload == "cfg5":
        return run_kcount(args, wl)
    if args.workload == "seal":
        return run_seal(args, wl)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_to_gpu_numa_node(local_rank, world) if world > 1 else "single rank: not pinned"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    paired = wl["paired"]
    L = wl["read_len"]

    # ---- table: built on rank 0, replicated by NCCL broadcast (no per-read collective afterwards) ----
    t_build = time.perf_counter()
    eng, stored, d_ref = build_engine(wl, rank, local_rank, world, lib, torch)
    t_build = time.perf_counter() - t_build
    if "stored" in wl:
        assert stored == wl["stored"], stored
    eng.set_max_read_len(L)

    # ---- HBM-resident batches from the device generator (distinct per rank and per buffer) ----------
    n_pairs = args.pairs_per_step
    n_reads = 2 * n_pairs
    nbuf = 2
    bufs = []
    for b in range(nbuf):
        d_bases = torch.empty(n_reads * L, dtype=torch.uint8, device=dev)
        d_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
        if args.workload == "cfg3":
            first = (rank * nbuf + b) * n_reads
            rc = lib.bbduk_b200_synth_contam(d_bases.data_ptr(), d_off.data_ptr(), n_reads, first, L, d_ref.data_ptr(),
                                             d_ref.numel(), C.c_uint64(1), 10, 100, 5, None)
        else:
            first = (rank * nbuf + b) * n_pairs
            rc = lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), n_pairs, first, L, C.c_uint64(1), 50, 5, None)
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
        eng.process_device(d_bases, d_off, n_reads, paired, outs, d_stats=d_stats, stream=stream.cuda_stream)

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

    # ---- the same step followed by trim-by-overlap (the full `... tpe tbo` command), timed separately ----
    tbo_info = None
    if args.workload == "cfg2":
        outs_t = dict(outs)
        outs_t["lo"] = torch.empty(n_reads, dtype=torch.int32, device=dev)
        d_tst = torch.zeros(2, dtype=torch.int64, device=dev)
        t_steps = max(2, min(args.steps, 5))

        def step_tbo(i):
            d_bases, d_off = bufs[i % nbuf]
            eng.process_device(d_bases, d_off, n_reads, True, outs_t, d_stats=d_stats, stream=stream.cuda_stream)
            eng.tbo_device(d_bases, None, d_off, n_reads, L, outs_t["lo"], outs_t["hi"], outs_t["flags"], None, d_tst,
                           stream=stream.cuda_stream)
        step_tbo(0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(t_steps):
                step_tbo(i + 1)
            e1.record(stream)
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tbo_info = {"reads_per_s": world * n_reads * t_steps / (float(tt.item()) * 1e-3), "ms_per_step": float(tt.item()) / t_steps,
                    "reads_trimmed_by_overlap_per_step": int(d_tst[0].item()) // (t_steps + 1)}

    # ---- the quality-trimming block (qtrim=rl trimq=10) on the same batch, timed alone ----
    qtrim_info = None
    if args.workload == "cfg2":
        # synthetic qualities: every read decays from Q40 towards its 3' end with its own slope, +-3 noise, clamped to [2,41]
        g = torch.Generator(device=dev)
        g.manual_seed(5 + rank)
        d_quals = torch.empty(n_reads * L, dtype=torch.uint8, device=dev)
        posn = torch.arange(L, device=dev, dtype=torch.int32)
        for c0 in range(0, n_reads, 1 << 20):
            c1 = min(n_reads, c0 + (1 << 20))
            slope = torch.randint(0, 45, (c1 - c0, 1), device=dev, dtype=torch.int32, generator=g)
            noise = torch.randint(-3, 4, (c1 - c0, L), device=dev, dtype=torch.int32, generator=g)
            d_quals[c0 * L:c1 * L] = (torch.clamp(40 - (posn * slope) // L + noise, 2, 41) + 33).to(torch.uint8).reshape(-1)
        del posn, slope, noise
        qcfg = eng.qtrim_cfg(qtrim_left=1, qtrim_right=1, trimq=10.0)
        d_qst = torch.zeros(8, dtype=torch.int64, device=dev)
        q_lo = torch.zeros(n_reads, dtype=torch.int32, device=dev)
        q_hi = torch.full((n_reads,), L, dtype=torch.int32, device=dev)
        q_fl = torch.zeros(n_reads, dtype=torch.uint8, device=dev)
        q_steps = max(3, min(args.steps, 10))

        def step_q(i):
            d_bases, d_off = bufs[i % nbuf]
            q_lo.zero_()
            q_hi.fill_(L)
            q_fl.zero_()
            eng.qtrim_device(d_bases, d_quals, d_off, n_reads, True, q_lo, q_hi, q_fl, qcfg, d_qst, stream=stream.cuda_stream)
        with torch.cuda.stream(stream):
            step_q(0)
        barrier()
        qe = [torch.cuda.Event(enable_timing=True) for _ in range(2 * q_steps)]
        with torch.cuda.stream(stream):
            for i in range(q_steps):
                d_bases, d_off = bufs[(i + 1) % nbuf]
                q_lo.zero_()
                q_hi.fill_(L)
                q_fl.zero_()
                qe[2 * i].record(stream)
                eng.qtrim_device(d_bases, d_quals, d_off, n_reads, True, q_lo, q_hi, q_fl, qcfg, d_qst, stream=stream.cuda_stream)
                qe[2 * i + 1].record(stream)
        barrier()
        q_ms = statistics.mean(qe[2 * i].elapsed_time(qe[2 * i + 1]) for i in range(q_steps))
        tq = torch.tensor([q_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        q_ms = float(tq.item())
        q_bytes = 2 * L + 4 + 9 + 9  # bases + qualities + offset in; lo, hi, flags in and out
        peak_q, _ = load_peak()
        qtrim_info = {"reads_per_s": world * n_reads / (q_ms * 1e-3), "ms_per_launch": q_ms, "algorithmic_bytes_per_read": q_bytes,
                      "hbm_frac": n_reads * q_bytes / (q_ms * 1e-3) / 1e9 / peak_q,
                      "reads_qtrimmed_per_launch": int(d_qst[0].item()) // (q_steps + 1)}
        del d_quals

    # ---- the low-entropy read filter (entropy=0.5) on the same batch, timed alone ----
    entropy_info = None
    if args.workload == "cfg2":
        ecfg = eng.entropy_cfg(cutoff=0.5)
        d_est = torch.zeros(2, dtype=torch.int64, device=dev)
        e_lo = torch.zeros(n_reads, dtype=torch.int32, device=dev)
        e_hi = torch.full((n_reads,), L, dtype=torch.int32, device=dev)
        e_fl = torch.zeros(n_reads, dtype=torch.uint8, device=dev)
        e_steps = 3
        with torch.cuda.stream(stream):
            eng.entropy_device(bufs[0][0], bufs[0][1], n_reads, True, e_lo, e_hi, e_fl, ecfg, d_est, stream=stream.cuda_stream)
        barrier()
        ee = [torch.cuda.Event(enable_timing=True) for _ in range(2 * e_steps)]
        with torch.cuda.stream(stream):
            for i in range(e_steps):
                d_bases, d_off = bufs[(i + 1) % nbuf]
                e_fl.zero_()
                ee[2 * i].record(stream)
                eng.entropy_device(d_bases, d_off, n_reads, True, e_lo, e_hi, e_fl, ecfg, d_est, stream=stream.cuda_stream)
                ee[2 * i + 1].record(stream)
        barrier()
        e_ms = statistics.mean(ee[2 * i].elapsed_time(ee[2 * i + 1]) for i in range(e_steps))
        te2 = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te2, op=dist.ReduceOp.MAX)
        e_ms = float(te2.item())
        entropy_info = {"reads_per_s": world * n_reads / (e_ms * 1e-3), "ms_per_launch": e_ms,
                        "reads_filtered_per_launch": int(d_est[0].item()) // (e_steps + 1)}

    # ---- end to end through the C ABI with pinned host buffers ------------------------------------
    e_pairs = args.e2e_pairs
    e_reads = 2 * e_pairs
    h_bases = torch.empty(e_reads * L, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(bufs[0][0][: e_reads * L])
    h_off = torch.arange(0, (e_reads + 1) * L, L, dtype=torch.int64).pin_memory()
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
        eng.process(hb, ho, paired, out=hout)
    barrier()
    x0 = eng.transfer_bytes()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        _, est = eng.process(hb, ho, paired, out=hout)
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    x1 = eng.transfer_bytes()
    # what really crossed PCIe per step, counted inside the library (packed chunks 0.375 B/base, ASCII chunks 1 B/base)
    h2d_wire, d2h_wire = (x1[0] - x0[0]) // e_steps, (x1[1] - x0[1]) // e_steps
    te = torch.tensor([e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_value = world * e_reads * e_steps / float(te.item())

    # ---- what bounds the host-buffer path on this box: the packer on all host threads and a plain pinned H2D copy ------------
    limits = None
    if args.workload == "cfg2" and rank == 0:
        from concurrent.futures import ThreadPoolExecutor
        nthr = max(1, (os.cpu_count() or 1) // max(1, world))
        nb_ = min(hb.size, 1 << 30) // (16 * 32 * nthr) * (16 * 32 * nthr)
        if nb_ > 0:
            gF = np.empty(nb_ // 16 + 64, np.uint32)
            gD = np.empty(nb_ // 16 + 64, np.uint16)
            sl = nb_ // nthr

            def one(i):
                lib.bbduk_b200_pack_bases(hb.ctypes.data + i * sl, sl, gF.ctypes.data + 4 * (i * sl // 16), gD.ctypes.data + 2 * (i * sl // 16))
            with ThreadPoolExecutor(nthr) as ex:
                list(ex.map(one, range(nthr)))
                t0 = time.perf_counter()
                list(ex.map(one, range(nthr)))
                pack_gbs = nb_ / (time.perf_counter() - t0) / 1e9
            d_tmp = torch.empty(nb_, dtype=torch.uint8, device=dev)
            d_tmp.copy_(h_bases[:nb_], non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d_tmp.copy_(h_bases[:nb_], non_blocking=True)
            torch.cuda.synchronize()
            h2d_gbs = nb_ / (time.perf_counter() - t0) / 1e9
            del d_tmp, gF, gD
            limits = {"host_pack_gb_per_s_ascii_in": round(pack_gbs, 1), "host_threads": nthr, "pinned_h2d_gb_per_s": round(h2d_gbs, 1),
                      "reads_per_s_if_all_chunks_packed": round(pack_gbs * 1e9 / L), "reads_per_s_if_all_chunks_ascii": round(h2d_gbs * 1e9 / (L + 8))}

    # ---- the same call for a host that already holds its reads 2-bit packed (bbduk_b200_process_packed): the stream is
    # packed once outside the timed region, as a packing FASTQ parser would hand it over ----------------------------------
    packed_info = None
    if args.workload == "cfg2":
        ng = (e_reads * L + 15) // 16
        h_F = torch.empty(ng + 16, dtype=torch.int32, pin_memory=True)
        h_D = torch.empty(ng + 32, dtype=torch.int16, pin_memory=True)
        assert lib.bbduk_b200_pack_bases(hb.ctypes.data, hb.size, h_F.data_ptr(), h_D.data_ptr()) == 0
        nF, nD = h_F.numpy().view(np.uint32), h_D.numpy().view(np.uint16)
        for _ in range(2):
            eng.process_packed(nF, nD, ho, paired, out=hout)
        barrier()
        y0 = eng.transfer_bytes()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            eng.process_packed(nF, nD, ho, paired, out=hout)
        torch.cuda.synchronize()
        p_dt = time.perf_counter() - t0
        y1 = eng.transfer_bytes()
        tp = torch.tensor([p_dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        packed_info = {"reads_per_s": world * e_reads * e_steps / float(tp.item()), "pairs_per_step_per_gpu": e_pairs,
                       "h2d_bytes_per_step": (y1[0] - y0[0]) // e_steps, "d2h_bytes_per_step": (y1[1] - y0[1]) // e_steps,
                       "note": "bbduk_b200_process_packed: 2-bit stream + defined bits packed by the caller (outside the timed region)"}
        del h_F, h_D

    # ---- the whole chain end to end (k-mer block + tbo + qtrim=rl trimq=10) through ONE C-ABI call, host buffers ----
    chain_info = chain_tbo_info = None
    if args.workload == "cfg2":
        c_pairs = min(e_pairs, 4 << 20)
        c_reads = 2 * c_pairs
        cb, co = hb[: c_reads * L], ho[: c_reads + 1]
        rngq = np.random.default_rng(3 + rank)
        h_cq = torch.empty(c_reads * L, dtype=torch.uint8, pin_memory=True)
        cq = h_cq.numpy()
        cq[:] = (33 + np.clip(40 - (np.arange(c_reads * L, dtype=np.int32) % L) * rngq.integers(0, 45, c_reads * L, dtype=np.int32) // L,
                              2, 41)).astype(np.uint8)
        tcfg, qcfg2 = eng.tbo_cfg(), eng.qtrim_cfg(qtrim_left=1, qtrim_right=1, trimq=10.0)
        # caller-owned pinned result arrays, as the JNI shim's staging has them: lo, hi, flags = 9 B per read come back
        cout = Outputs(0)
        cout.n = c_reads
        cpin = {"lo": torch.empty(c_reads, dtype=torch.int32, pin_memory=True),
                "hi": torch.empty(c_reads, dtype=torch.int32, pin_memory=True),
                "flags": torch.empty(c_reads, dtype=torch.uint8, pin_memory=True)}
        cout.lo, cout.hi, cout.flags = (cpin[k].numpy() for k in ("lo", "hi", "flags"))
        cout.id0 = cout.id0b = cout.count = None
        eng.process_chain(cb, cq, co, True, tbo=tcfg, qtrim=qcfg2, out=cout)
        barrier()
        z0 = eng.transfer_bytes()
        t0 = time.perf_counter()
        c_steps = 3
        for _ in range(c_steps):
            _, _, ct2, cq8, _ = eng.process_chain(cb, cq, co, True, tbo=tcfg, qtrim=qcfg2, out=cout)
        c_dt = time.perf_counter() - t0
        z1 = eng.transfer_bytes()
        tc = torch.tensor([c_dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        # exactly BASELINE.json configs[1] (`ktrim=r k=23 mink=11 hdist=1 tpe tbo`): k-mer block + trim by overlap, no qualities
        eng.process_chain(cb, None, co, True, tbo=tcfg, out=cout)
        barrier()
        z2 = eng.transfer_bytes()
        t0 = time.perf_counter()
        for _ in range(c_steps):
            _, _, ct2b, _, _ = eng.process_chain(cb, None, co, True, tbo=tcfg, out=cout)
        cb_dt = time.perf_counter() - t0
        z3 = eng.transfer_bytes()
        tcb = torch.tensor([cb_dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tcb, op=dist.ReduceOp.MAX)
        chain_tbo_info = {"reads_per_s": world * c_reads * c_steps / float(tcb.item()), "pairs_per_call_per_gpu": c_pairs,
                          "h2d_bytes_per_call": int((z3[0] - z2[0]) // c_steps), "host_input_bytes_per_call": int(cb.nbytes + 8 * (c_reads + 1)),
                          "d2h_bytes_per_call": 9 * c_reads, "reads_trimmed_by_overlap": int(ct2b[0]),
                          "note": "h2d = bytes that crossed PCIe as counted by the library (chunks of A C G T N cross 2-bit packed and are spelled out again on the device for tbo)"}
        chain_info = {"reads_per_s": world * c_reads * c_steps / float(tc.item()), "pairs_per_call_per_gpu": c_pairs,
                      "h2d_bytes_per_call": int((z1[0] - z0[0]) // c_steps), "host_input_bytes_per_call": int(cb.nbytes + cq.nbytes + 8 * (c_reads + 1)),
                      "d2h_bytes_per_call": 9 * c_reads, "reads_trimmed_by_overlap": int(ct2[0]), "reads_qtrimmed": int(cq8[0])}

    # ---- sanity: EVERY rank checks a slice of its own timed batch against the CPU oracle (checker only); the flag in the
    # line is the AND over ranks, so a rank whose replicated table arrived broken cannot hide behind rank 0 -----------
    parity = None
    o = None
    if args.workload == "cfg2" or args.verify:
        from oracle.oracle import Oracle
        chk = 20000
        o = Oracle(make_cfg(**wl["cfg"]))
        if d_ref is None:
            rb, roff = adapters_ref()
        else:
            rb = d_ref.cpu().numpy()
            roff = np.arange(0, rb.size + 1, wl["ref"][1], dtype=np.int64)
        o.add_ref(rb, roff)
        stored_o = o.finalize()
        step(args.warmup + args.steps - 1)
        torch.cuda.synchronize()
        db, _ = bufs[(args.warmup + args.steps - 1) % nbuf]
        hb2 = db[: 2 * chk * L].cpu().numpy()  # the device generators are tested against synth.py byte for byte
        ho2 = np.arange(0, (2 * chk + 1) * L, L, dtype=np.int64)
        want, _ = o.process(hb2, ho2, paired, threads=max(1, min(8, (os.cpu_count() or 1) // world)))
        mine = bool(np.array_equal(outs["hi"][: 2 * chk].cpu().numpy(), want.hi) and
                    np.array_equal(outs["id0"][: 2 * chk].cpu().numpy(), want.id0) and
                    np.array_equal(outs["flags"][: 2 * chk].cpu().numpy(), want.flags) and stored_o == stored)
        # the replicated table itself: an order-independent checksum of (key, id) over all slots must equal rank 0's
        csum, n_keys = table_checksum(eng, torch)
        flag = torch.tensor([1 if mine else 0, csum, -csum, n_keys, -n_keys], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        f = flag.cpu().tolist()
        tables_equal = (f[1] == -f[2]) and (f[3] == -f[4])  # min == max over ranks
        parity = bool(f[0] == 1 and tables_equal)
        parity_detail = {"ranks_checked": world, "reads_checked_per_rank": 2 * chk, "all_ranks_match_oracle": bool(f[0] == 1),
                         "table_checksum_equal_on_all_ranks": bool(tables_equal), "table_keys": int(f[3])}
    else:
        parity_detail = None

    if rank == 0:
        peak, peak_kind = load_peak()
        kern_ms = statistics.mean(step_ms)
        achieved = n_reads * wl["alg_bytes"] / (kern_ms * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        cpu_pairs = args.cpu_pairs
        cpu_rate = None
        if world == 1 and cpu_pairs > 0 and o is not None:
            # the sample is the head of the timed workload, copied back from the device generator; the oracle (CPU restatement,
            # all host cores) is the one the parity check above built (cfg 3 / 4: only with --verify, its table takes minutes)
            cpu_pairs = min(cpu_pairs, n_pairs)
            cb = bufs[0][0][: 2 * cpu_pairs * L].cpu().numpy()
            co = np.arange(0, (2 * cpu_pairs + 1) * L, L, dtype=np.int64)
            o.process(cb[: 4000 * L], co[:4001], paired, threads=cores)
            t0 = time.perf_counter()
            o.process(cb, co, paired, threads=cores)
            cpu_rate = 2 * cpu_pairs / (time.perf_counter() - t0)
        line = {
            "metric": "bbduk_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": wl["desc"], "pairs_per_step_per_gpu": n_pairs, "read_len": L,
                       "stored_kmers": stored, "l2": f"inputs {n_reads * L / 2**20:.0f} MiB per step > 126 MB L2, "
                       f"{nbuf} alternating buffers, no flush", "table_build_s": round(t_build, 3),
                       "parity_vs_oracle_on_timed_batch": parity, "parity_detail": parity_detail,
                       "kmer_block_plus_tbo": tbo_info,
                       "qtrim_block": qtrim_info, "entropy_block": entropy_info,
                       "chain_e2e_kmer_tbo": chain_tbo_info, "chain_e2e_kmer_tbo_qtrim": chain_info, "e2e_packed_input": packed_info, "e2e_host_limits": limits, "host_numa_pinning": numa},
            "e2e": {"value": e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d_wire), "d2h_bytes_per_step": int(d2h_wire),
                    "host_input_bytes_per_step": h2d, "host_output_bytes_per_step": d2h,
                    "pairs_per_step_per_gpu": e_pairs, "steps": e_steps,
                    "note": "ASCII bases + int64 offsets in pinned host memory in, id0/hi/flags out, through bbduk_b200_process; "
                            "h2d/d2h = bytes that crossed PCIe as counted by the library (most chunks cross 2-bit packed)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (NCU_TRAFFIC_BYTES_PER_READ[args.workload] * n_reads
                                     if args.workload in NCU_TRAFFIC_BYTES_PER_READ else None),
                         "traffic_unit": "bytes per launch (ncu dram read+write per read x reads per launch)",
                         "traffic_source": TRAFFIC_SOURCE.get(args.workload),
                         "peak_source": peak_kind, "kernel": wl["kernel"],
                         "algorithmic_bytes_per_read": wl["alg_bytes"], "ms_per_launch": kern_ms},
        }
        if args.workload in ("cfg3", "cfg4"):
            # DRAM cannot fetch less than a 32-byte sector per probe: the sector-granular variant of SURVEY.md 8d
            sect = L + 12 + 32 * (wl["alg_bytes"] - L - 12) // 8
            line["roofline"]["sector_granular_bytes_per_read"] = sect
            line["roofline"]["sector_granular_frac"] = n_reads * sect / (kern_ms * 1e-3) / 1e9 / peak
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": "reads/s", "cores": cores, "kind": "port",
                                    "sample": f"{2 * cpu_pairs} reads of the same synthetic workload (seed 1), oracle C port, "
                                              f"{cores} threads"}
        emit(json.dumps(line))
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
    ap.add_argument("--e2e-pairs", type=int, default=0, help="pairs per GPU per end-to-end step (host buffers); 0 = the same batch as --pairs-per-step")
    ap.add_argument("--cpu-pairs", type=int, default=4 << 20, help="pairs of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--ref-pairs", type=int, default=1 << 19, help="pairs per step of --impl reference")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="cfg2 = the headline line (default); cfg3/cfg4 = HBM-resident tables; cfg5 = kmercountexact; seal = Seal's matching block")
    ap.add_argument("--verify", action="store_true", help="cfg3/cfg4: also build the CPU oracle's table and check a slice")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.e2e_pairs <= 0:
        args.e2e_pairs = args.pairs_per_step if args.workload != "cfg5" else min(args.pairs_per_step, 2 << 20)
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
