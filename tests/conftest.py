"""Shared fixtures.  `-m "not gpu"`: oracle vs golden vectors, host logic, ABI export check.
`-m gpu`: parity tests proper, all through the C ABI of libnb200.so."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle():
    o = entry.load_oracle()
    o.build(ref=False)
    return o


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def lib(pkg):
    """libnb200.so, built on demand (nvcc cross-compiles without a GPU)."""
    if not os.path.exists(pkg._lib.LIB_PATH):
        pkg._lib.build()
    return pkg._lib.load()
