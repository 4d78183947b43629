"""pytest configuration: the `gpu` marker, import paths, and a lazily built C-ABI library."""
import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
PKG = REPO / "hpc-ops_b200"
for p in (str(REPO), str(PKG), str(Path(__file__).resolve().parent)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    so = PKG / "hpc" / "_C.so"
    if not so.exists():
        sys.path.insert(0, str(PKG))
        import build as _b  # hpc-ops_b200/build.py

        _b.build()
    return so


@pytest.fixture(scope="session")
def lib_path():
    return _ensure_built()


@pytest.fixture(scope="session")
def hpc(lib_path):
    import hpc as _hpc

    return _hpc


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
