import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    """liboracle.so, (re)built from oracle/*.c when needed (gcc only; no GPU)."""
    import oracle_py
    return oracle_py.load()


@pytest.fixture(scope="session")
def emd():
    """The product library; built here when missing (nvcc cross-compiles without a GPU)."""
    import examinimd_b200
    if not examinimd_b200.LIB_PATH.exists():
        from examinimd_b200 import build
        build.build()
    examinimd_b200.lib()
    return examinimd_b200
