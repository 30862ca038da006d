import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _library_built():
    """The data-format tests (TFRecord CRC, WAV decode, TF checkpoints) call HOST entry points of the shared library:
    build it when a fresh checkout has none (an existing one is never rebuilt here)."""
    from gansynth_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture
def emu():
    """Product host code on the torch-CPU emulation of the kernel API (tests/emu_backend.py)."""
    import gansynth_b200.functional as F
    import gansynth_b200.models as M
    import gansynth_b200.ops as ops
    from emu_backend import EmuBackend
    prev = F.K
    F.set_backend(EmuBackend())
    store = ops.set_default_store(ops.VariableStore(device="cpu", seed=0))
    M.reset_global_step()
    yield store
    F.set_backend(prev)
    ops.set_default_store(None)
    M.reset_global_step()


@pytest.fixture
def cuda_store():
    import torch
    import gansynth_b200.models as M
    import gansynth_b200.ops as ops
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    store = ops.set_default_store(ops.VariableStore(device="cuda", seed=0))
    M.reset_global_step()
    yield store
    ops.set_default_store(None)
    M.reset_global_step()
