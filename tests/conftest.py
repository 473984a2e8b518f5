import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def adapters():
    from bbtools_b200.fasta import read_fasta
    names, bases, offsets = read_fasta(os.path.join(GOLDEN, "adapters.fa"))
    return names, bases, offsets


@pytest.fixture(scope="session")
def adapter_seqs(adapters):
    names, bases, offsets = adapters
    return [bytes(bases[offsets[i]:offsets[i + 1]]).decode() for i in range(len(names))]
