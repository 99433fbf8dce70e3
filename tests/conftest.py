import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


class _Tuning:
    """environment knobs of the lattice kernels: the library reads them once, so every change is followed by
    tsu_ising2d_reload_tuning()"""

    def __init__(self, monkeypatch):
        self.mp = monkeypatch

    def _reload(self):
        from tsu_emulator_b200 import _lib

        _lib.load().tsu_ising2d_reload_tuning()

    def setenv(self, name, value):
        self.mp.setenv(name, value)
        self._reload()

    def delenv(self, name):
        self.mp.delenv(name, raising=False)
        self._reload()


@pytest.fixture
def tuning(monkeypatch):
    t = _Tuning(monkeypatch)
    yield t
    monkeypatch.undo()
    t._reload()
