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


@pytest.fixture(params=['emulator', pytest.param('cuda', marks=pytest.mark.gpu)])
def kdev(request, monkeypatch):
    """Device the reference-facing API tests run on: the CPU emulator of the C ABI (``-m "not gpu"``) or the real
    library on cuda:0 (``-m gpu``)."""
    import torch
    if request.param == 'emulator':
        import cabi_emulator
        from exploring_meta_b200 import _lib, engine
        monkeypatch.setattr(_lib, '_lib', cabi_emulator.EmulatedLib())
        monkeypatch.setattr(engine, '_require_cuda', lambda device: None)
        return torch.device('cpu')
    return torch.device('cuda', 0)
