import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodal-rssm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_ready():
    """A CUDA device AND the in-tree sm_100a library (the product has no fallback path)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
        from mrssm_b200 import _lib
        _lib.load()
        return True, ""
    except Exception as e:            # library missing / not loadable
        return False, f"libmrssm_b200.so not loadable: {e}"


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a GPU or without the built library, so a plain `pytest`
    works anywhere; on the B200 box `-m gpu` runs them all."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu") is not None]
    if not gpu_items:
        return
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason=f"needs the B200 path: {why}")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
