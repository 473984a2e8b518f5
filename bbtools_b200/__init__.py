"""bbtools_b200: B200-native BBDuk k-mer match-and-trim hot path (CUDA behind a C ABI).

The package holds only what the path needs: csrc/ (CUDA kernels + the C ABI of
include/bbduk_b200.h), the ctypes binding, and the host-side mirror of jgi.BBDuk's
read-in/read-out surface (bbduk.py), of KmerCountExact's counting table (kcount.py) and of Seal's
loader + matching block (seal.py). There is no CPU fallback: importing the binding fails loudly when
libbbduk_b200.so has not been built.
"""
from ._abi import (BBDukCfg, BBDukOut, BBDukStats, BBDukTboCfg, F_DISCARDED, F_KTRIMMED, F_REMOVED, F_SPLIT, F_TBO, F_TPE, GEN_JGI, GEN_S,
                   Outputs, default_cfg, make_cfg)

__all__ = ["BBDukCfg", "BBDukOut", "BBDukStats", "Outputs", "default_cfg", "make_cfg", "GEN_JGI", "GEN_S",
           "F_DISCARDED", "F_REMOVED", "F_KTRIMMED", "F_TPE", "F_SPLIT", "F_TBO", "BBDukTboCfg"]
