import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) on a machine without a CUDA device."""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the GPU box with -m gpu)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
