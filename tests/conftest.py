import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (CUDA library by cross-compilation, oracle, emulator)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def emu_ctx(built):
    """The kernels compiled for the CPU SIMT emulator -- test tier only, never product."""
    from biogo_b200 import _lib
    import helpers

    return _lib.Context(0, helpers.EMU_LIB)


@pytest.fixture(scope="session")
def gpu_ctx():
    from biogo_b200 import _lib

    return _lib.Context(0)  # fails loudly if the CUDA library or the GPU is missing
