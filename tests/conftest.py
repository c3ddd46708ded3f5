import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """the CPU oracle module (test infrastructure); builds oracle/liboracle.so on first use"""
    from oracle import oracle as O
    O.lib("port")
    return O


@pytest.fixture(scope="session")
def have_ref(oracle):
    return oracle.have_reference()


def per_query_sets(offsets, counts, cand):
    """list of sorted candidate arrays, one per query"""
    offsets = np.asarray(offsets)
    counts = np.asarray(counts)
    cand = np.asarray(cand)
    return [np.sort(cand[o:o + c]) for o, c in zip(offsets, counts)]
