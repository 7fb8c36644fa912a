import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.join(ROOT, "tests")
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def emu_lib():
    """CPU execution-model emulation of the kernel sources (tests/emu) bound through the same ctypes
    layer as the product library.  TEST-ONLY: lets the no-GPU suite run the real kernel source."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("emulation tests run only where no GPU exists (the real library is tested with -m gpu)")
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    path = build_emu.build()
    from dna_b200 import _lib
    _lib._use_library_for_tests(path)
    return _lib.lib()


def pytest_collection_modifyitems(config, items):
    # GPU tests bind the real library; emulation tests rebind the ctypes layer. Keep them apart.
    import torch
    if not torch.cuda.is_available():
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)
