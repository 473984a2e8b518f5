"""Host-side mirror of KmerCountExact's counting surface over include/kcount_b200.h.

`KmerTableSetGPU` plays the role of kmer.KmerTableSet for this path (kmer/KmerTableSet.java:400-716):
add reads, then read `Unique Kmers`, the k-mer count histogram (khist=) and the (key,count) dump (out=).
Multi-GPU: every rank counts its own slice of the reads into a private table; ONE exchange
at the end re-partitions the entries by owner = mix(key) % world (all_to_all over NCCL) and merges them,
after which every key lives on exactly one rank, so `Unique Kmers` and the histogram are plain sums.
No CPU fallback: the CUDA library must be present."""
import ctypes as C

import numpy as np

from . import _lib

HISTMAX_DEFAULT = 100000  # jgi/KmerCountExact.java:1068


class KmerTableSetGPU:
    def __init__(self, k=31, rcomp=True, initial_keys=0, device=-1):
        self.lib = _lib.load()
        self.k = k
        self.rcomp = bool(rcomp)
        self.device = device
        h = C.c_void_p()
        rc = self.lib.kcount_b200_create(k, 1 if rcomp else 0, int(initial_keys), device, C.byref(h))
        if rc:
            raise RuntimeError(self.lib.kcount_b200_last_error(None).decode())
        self.h = h

    def _ck(self, rc):
        if rc:
            raise RuntimeError(self.lib.kcount_b200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.kcount_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- kmer/KmerTableSet.java:652-716 for a batch ---------------------------------------------
    def add_reads(self, bases: np.ndarray, offsets: np.ndarray):
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        self._ck(self.lib.kcount_b200_add_reads(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1))

    def add_reads_device(self, d_bases, d_offsets, n_reads, n_bases, stream=None):
        """device tensors (uint8 bases, int32/uint32 offsets[n_reads+1])"""
        self._ck(self.lib.kcount_b200_add_reads_device(self.h, d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, n_bases,
                                                       stream))

    def stats(self):
        v = np.zeros(4, np.int64)
        self._ck(self.lib.kcount_b200_stats(self.h, v.ctypes.data))
        return {"reads_in": int(v[0]), "bases_in": int(v[1]), "kmers_in": int(v[2]), "unique_kmers": int(v[3])}

    def khist(self, histmax=HISTMAX_DEFAULT):
        hist = np.zeros(histmax + 1, np.int64)
        self._ck(self.lib.kcount_b200_khist(self.h, histmax, hist.ctypes.data))
        return hist

    def dump(self, mincount=1, maxcount=0x7FFFFFFF):
        n = C.c_int64()
        cap = self.stats()["unique_kmers"]
        keys = np.zeros(max(cap, 1), np.uint64)
        counts = np.zeros(max(cap, 1), np.int32)
        self._ck(self.lib.kcount_b200_dump(self.h, mincount, maxcount, keys.ctypes.data, counts.ctypes.data, cap, C.byref(n)))
        m = min(int(n.value), cap)
        return keys[:m], counts[:m]

    def table_info(self):
        v = np.zeros(3, np.int64)
        self._ck(self.lib.kcount_b200_table_info(self.h, v.ctypes.data))
        return {"n_slots": int(v[0]), "bytes": int(v[1]), "launches": int(v[2])}

    # ---- the exchange's two device-side halves ------------------------------------------------------
    def export_partitioned(self, n_parts):
        """-> (keys int64 tensor, counts int32 tensor, sizes list) on this handle's device, grouped by owner"""
        import torch
        n = self.stats()["unique_kmers"]
        dev = torch.device("cuda", torch.cuda.current_device() if self.device < 0 else self.device)
        keys = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        counts = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        sizes = np.zeros(n_parts, np.int64)
        torch.cuda.synchronize()
        self._ck(self.lib.kcount_b200_export_partitioned(self.h, n_parts, keys.data_ptr(), counts.data_ptr(),
                                                         sizes.ctypes.data, None))
        return keys[:n], counts[:n], [int(x) for x in sizes]

    def merge(self, keys, counts):
        import torch
        torch.cuda.synchronize()
        self._ck(self.lib.kcount_b200_merge_device(self.h, keys.data_ptr(), counts.data_ptr(), int(keys.numel()), None))
        torch.cuda.synchronize()

    def new_like(self, initial_keys=0):
        return KmerTableSetGPU(self.k, self.rcomp, initial_keys, self.device)


def exchange_counts(table, group=None):
    """The ONE collective of config 5: re-partition a rank-private table by key owner.

    `table` needs export_partitioned(n) -> (keys, counts, sizes), new_like(initial_keys) and merge(keys, counts);
    KmerTableSetGPU on GPUs (NCCL); the gloo tests pass a CPU stand-in with the same three methods.
    Returns the rank's owner table: afterwards every key is on exactly one rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    keys, counts, sizes = table.export_partitioned(world)
    dev = keys.device
    send = torch.tensor(sizes, dtype=torch.int64, device=dev)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    rsz = [int(x) for x in recv.tolist()]
    rkeys = torch.empty(sum(rsz), dtype=keys.dtype, device=dev)
    rcounts = torch.empty(sum(rsz), dtype=counts.dtype, device=dev)
    dist.all_to_all_single(rkeys, keys.contiguous(), output_split_sizes=rsz, input_split_sizes=sizes, group=group)
    dist.all_to_all_single(rcounts, counts.contiguous(), output_split_sizes=rsz, input_split_sizes=sizes, group=group)
    owner = table.new_like(initial_keys=sum(rsz))
    owner.merge(rkeys, rcounts)
    return owner


def global_summary(owner_table, histmax=HISTMAX_DEFAULT, group=None):
    """Unique k-mers and khist over all ranks after exchange_counts (sums; every key has one owner)."""
    import torch
    import torch.distributed as dist

    hist = torch.from_numpy(np.asarray(owner_table.khist(histmax), np.int64).copy())
    uniq = torch.tensor([owner_table.stats()["unique_kmers"]], dtype=torch.int64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            hist, uniq = hist.cuda(), uniq.cuda()
        dist.all_reduce(hist, group=group)
        dist.all_reduce(uniq, group=group)
    return int(uniq.item()), hist.cpu().numpy()


def write_khist(path, hist, print_zeros=False):
    """khist= file format of kmer/AbstractKmerTableSet.java:563-638 (2 columns, header on)."""
    with open(path, "w") as f:
        f.write("#Depth\tCount\n")
        for d in range(1, len(hist)):
            if print_zeros or hist[d] > 0:
                f.write(f"{d}\t{int(hist[d])}\n")
