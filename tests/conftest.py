import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()  # builds oracle/_build/liboracle.so on first use
    return oracle


@pytest.fixture(scope="session")
def native_lib():
    """The in-tree CUDA library; (re)built when stale.  Fails loudly when it cannot be built."""
    from metricsfm_b200 import build, _lib
    build.build_native()
    return _lib.load()


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["basic", "ragged", "ties", "small_gate", "min_ok"]


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return {name: dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"))) for name in GOLDEN_CASES + ["float_unit"]}
