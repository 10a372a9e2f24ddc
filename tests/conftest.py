"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` runs on the CPU-only build box (oracle vs golden vectors, host logic, C-ABI symbol check);
`-m gpu` runs on a B200 and compares the CUDA path (through the C ABI) with the oracle bit-for-bit.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def _ensure_built():
    """Build the C-ABI library in-tree when it is missing or stale (nvcc cross-compiles without a GPU)."""
    builder = importlib.import_module("mpc-protocols_b200.build")
    if builder.needs_build():
        builder.build()


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    _ensure_built()


@pytest.fixture(scope="session")
def hb():
    """The product package (directory name has a hyphen, so it is imported by string)."""
    _ensure_built()
    return importlib.import_module("mpc-protocols_b200")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import cmodel

    cmodel.load()
    return cmodel


@pytest.fixture(scope="session")
def ctx(hb):
    c = hb.Context(0)
    yield c
    c.close()
