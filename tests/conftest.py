"""pytest configuration: the `gpu` marker and shared imports.

CPU suite  (`-m "not gpu"`): oracle vs analytic known answers and golden vectors, asset decoding, host logic,
                              C-ABI export check.  No kernel runs.
GPU suite  (`-m gpu`)       : parity of the CUDA path (through the C ABI) against the oracle.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gpx():
    """The product binding (c-game-engine_b200).  Importing never touches the oracle."""
    return importlib.import_module("c-game-engine_b200")


@pytest.fixture(scope="session")
def scenes():
    return importlib.import_module("c-game-engine_b200.scenes")


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    _orc.lib()
    return _orc
