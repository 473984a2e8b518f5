"""ctypes loader of libbbduk_b200.so. No fallback: a missing library is an ImportError with build advice."""
import ctypes as C
import os

from ._abi import ABI_VERSION, BBDukCfg, BBDukChainCfg, BBDukEntropyCfg, BBDukOut, BBDukQtrimCfg, BBDukStats, BBDukTableDesc, BBDukTboCfg
from .seal import SealCfg, SealOut, SealStats

_HERE = os.path.dirname(os.path.abspath(__file__))
# BBDUK_B200_LIB: another build of the same library (e.g. the -DBB_FAST_COUNT debug build, `make -C bbtools_b200/csrc debug`)
LIB_PATH = os.environ.get("BBDUK_B200_LIB") or os.path.join(_HERE, "libbbduk_b200.so")

# every symbol include/bbduk_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("bbduk_b200_version", C.c_int, []),
    ("bbduk_b200_cfg_default", None, [C.POINTER(BBDukCfg)]),
    ("bbduk_b200_describe_cfg", C.c_int, [C.POINTER(BBDukCfg), C.c_void_p]),
    ("bbduk_b200_create", C.c_int, [C.POINTER(BBDukCfg), C.POINTER(C.c_void_p)]),
    ("bbduk_b200_add_ref", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    ("bbduk_b200_finalize", C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    ("bbduk_b200_process", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                     C.POINTER(BBDukOut), C.POINTER(BBDukStats)]),
    ("bbduk_b200_process_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                            C.POINTER(BBDukOut), C.c_void_p, C.c_void_p]),
    ("bbduk_b200_set_max_read_len", C.c_int, [C.c_void_p, C.c_int32]),
    ("bbduk_b200_scaffold_counts", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    ("bbduk_b200_table_describe", C.c_int, [C.c_void_p, C.POINTER(BBDukTableDesc)]),
    ("bbduk_b200_table_alloc", C.c_int, [C.c_void_p, C.POINTER(BBDukTableDesc)]),
    ("bbduk_b200_table_commit", C.c_int, [C.c_void_p]),
    ("bbduk_b200_replicate", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    ("bbduk_b200_replica_transport", C.c_int, [C.c_void_p]),
    ("bbduk_b200_process_packed", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                            C.POINTER(BBDukOut), C.POINTER(BBDukStats)]),
    ("bbduk_b200_process_sharded", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                             C.POINTER(BBDukOut), C.POINTER(BBDukStats)]),
    ("bbduk_b200_scaffold_counts_sum", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    ("bbduk_b200_table_export", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    ("bbduk_b200_ref_kmers", C.c_int64, [C.c_void_p]),
    ("bbduk_b200_transfer_bytes", C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("bbduk_b200_launch_count", C.c_int64, [C.c_void_p]),
    ("bbduk_b200_last_error", C.c_char_p, [C.c_void_p]),
    ("bbduk_b200_destroy", None, [C.c_void_p]),
    ("bbduk_b200_synth_pairs", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_uint64,
                                         C.c_int32, C.c_int32, C.c_void_p]),
    ("bbduk_b200_tbo_cfg_default", None, [C.POINTER(BBDukTboCfg)]),
    ("bbduk_b200_tbo", C.c_int, [C.c_void_p, C.POINTER(BBDukTboCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_tbo_device", C.c_int, [C.c_void_p, C.POINTER(BBDukTboCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_qtrim_cfg_default", None, [C.POINTER(BBDukQtrimCfg)]),
    ("bbduk_b200_qtrim", C.c_int, [C.c_void_p, C.POINTER(BBDukQtrimCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_qtrim_device", C.c_int, [C.c_void_p, C.POINTER(BBDukQtrimCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_entropy_cfg_default", None, [C.POINTER(BBDukEntropyCfg)]),
    ("bbduk_b200_entropy", C.c_int, [C.c_void_p, C.POINTER(BBDukEntropyCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_entropy_mask", C.c_int, [C.c_void_p, C.POINTER(BBDukEntropyCfg), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_entropy_mask_device", C.c_int, [C.c_void_p, C.POINTER(BBDukEntropyCfg), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p]),
    ("bbduk_b200_entropy_device", C.c_int, [C.c_void_p, C.POINTER(BBDukEntropyCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_chain_cfg_default", None, [C.POINTER(BBDukChainCfg)]),
    ("bbduk_b200_process_chain", C.c_int, [C.c_void_p, C.POINTER(BBDukChainCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_int32, C.POINTER(BBDukOut), C.POINTER(BBDukStats), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_pack_bases", C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ("bbduk_b200_synth_reference", C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]),
    ("bbduk_b200_synth_contam", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                          C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    # include/kcount_b200.h
    ("kcount_b200_create", C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]),
    ("kcount_b200_add_reads", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    ("kcount_b200_add_reads_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    ("kcount_b200_stats", C.c_int, [C.c_void_p, C.c_void_p]),
    ("kcount_b200_khist", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    ("kcount_b200_dump", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.POINTER(C.c_int64)]),
    ("kcount_b200_export_partitioned", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("kcount_b200_merge_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("kcount_b200_table_info", C.c_int, [C.c_void_p, C.c_void_p]),
    ("kcount_b200_synth_reads", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64,
                                          C.c_uint64, C.c_int32, C.c_void_p]),
    ("kcount_b200_last_error", C.c_char_p, [C.c_void_p]),
    ("kcount_b200_destroy", None, [C.c_void_p]),
    # include/seal_b200.h
    ("seal_b200_cfg_default", None, [C.POINTER(SealCfg)]),
    ("seal_b200_create", C.c_int, [C.POINTER(SealCfg), C.POINTER(C.c_void_p)]),
    ("seal_b200_add_ref", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    ("seal_b200_finalize", C.c_int, [C.c_void_p, C.c_void_p]),
    ("seal_b200_n_units", C.c_int64, [C.c_void_p, C.c_int64, C.c_int32]),
    ("seal_b200_process", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.POINTER(SealOut),
                                    C.POINTER(SealStats)]),
    ("seal_b200_process_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.POINTER(SealOut),
                                           C.c_void_p, C.c_void_p]),
    ("seal_b200_scaffold_counts", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    ("seal_b200_table_export", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    ("seal_b200_launch_count", C.c_int64, [C.c_void_p]),
    ("seal_b200_last_error", C.c_char_p, [C.c_void_p]),
    ("seal_b200_destroy", None, [C.c_void_p]),
    # include/fastq_b200.h
    ("fastq_b200_index", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32]),
    ("fastq_b200_gather", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]),
    ("fastq_b200_format", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_int64), C.c_int32]),
]

_LIB = None


def load():
    """Load the CUDA library; raises ImportError (never falls back) if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C bbtools_b200/csrc`). bbtools_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    v = lib.bbduk_b200_version()
    if v != ABI_VERSION:
        raise ImportError(f"libbbduk_b200.so ABI version {v} != binding {ABI_VERSION}")
    _LIB = lib
    return lib
