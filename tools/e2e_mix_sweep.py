#!/usr/bin/env python
"""How should the chunks of bbduk_b200_process cross PCIe when N ranks share one host?  Run under torchrun (or alone):
every rank opens one handle per setting of the ASCII / packed chunk mix (BBDUK_B200_ASCII_EVERY, BBDUK_B200_PCIE_GBS are
read when the handle is opened), times the same pinned cfg-2 batch through it and rank 0 prints whole-job reads/s.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_mix_sweep.py [--pairs N]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4 << 20)
    ap.add_argument("--steps", type=int, default=4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench
    from bbtools_b200 import _lib, make_cfg
    from bbtools_b200._abi import Outputs
    from bbtools_b200.bbduk import BBDukIndexGPU
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        bench.pin_to_gpu_numa_node(local_rank, world)
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    L = 150
    n_reads = 2 * args.pairs
    d_bases = torch.empty(n_reads * L, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n_reads + 1, dtype=torch.int32, device=dev)
    assert lib.bbduk_b200_synth_pairs(d_bases.data_ptr(), d_off.data_ptr(), args.pairs, rank * args.pairs, L, C.c_uint64(1), 50, 5, None) == 0
    h_bases = torch.empty(n_reads * L, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d_bases)
    h_off = torch.arange(0, (n_reads + 1) * L, L, dtype=torch.int64).pin_memory()
    del d_bases, d_off
    hout = Outputs(0)
    hout.n = n_reads
    pinned = {"id0": torch.empty(n_reads, dtype=torch.int32, pin_memory=True), "hi": torch.empty(n_reads, dtype=torch.int32, pin_memory=True),
              "flags": torch.empty(n_reads, dtype=torch.uint8, pin_memory=True)}
    hout.id0, hout.hi, hout.flags = (pinned[k].numpy() for k in ("id0", "hi", "flags"))
    hout.id0b = hout.lo = hout.count = None
    hb, ho = h_bases.numpy(), h_off.numpy()
    rb, roff = bench.adapters_ref()
    settings = [("adaptive (default)", {}), ("all ASCII", {"BBDUK_B200_ASCII_EVERY": "1"}), ("all packed", {"BBDUK_B200_ASCII_EVERY": "0"}),
                ("every 2nd ASCII", {"BBDUK_B200_ASCII_EVERY": "2"}), ("every 4th ASCII", {"BBDUK_B200_ASCII_EVERY": "4"})]
    ref_hi = None
    for name, env in settings:
        for k in ("BBDUK_B200_ASCII_EVERY", "BBDUK_B200_PCIE_GBS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        eng = BBDukIndexGPU(make_cfg(device=local_rank, k=23, mink=11, hdist=1, ktrim_right=1, trim_pairs_evenly=1))
        eng.add_ref(rb, roff)
        eng.finalize()
        eng.set_max_read_len(L)
        for _ in range(2):
            eng.process(hb, ho, True, out=hout)
        if ref_hi is None:
            ref_hi = pinned["hi"].clone()
        same = bool(torch.equal(ref_hi, pinned["hi"]))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        x0 = eng.transfer_bytes()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.process(hb, ho, True, out=hout)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        x1 = eng.transfer_bytes()
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"setting": name, "n_gpus": world, "reads_per_s": world * n_reads * args.steps / float(te.item()),
                              "h2d_bytes_per_read": (x1[0] - x0[0]) / args.steps / n_reads, "results_equal": same}), flush=True)
        del eng
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
