import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture
def emulated_lib(monkeypatch):
    """Routes the product's host-side launch programs to the CPU emulator of the C ABI
    (tests/cabi_emulator.py) so the orchestration can be checked without a GPU."""
    import cabi_emulator
    from exploring_meta_b200 import _lib, engine
    lib = cabi_emulator.EmulatedLib()
    monkeypatch.setattr(_lib, '_lib', lib)
    monkeypatch.setattr(engine, '_require_cuda', lambda device: None)
    return lib
