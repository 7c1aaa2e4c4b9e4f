import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The shared library is a build artefact (git-ignored): on a fresh checkout build it in-tree once (nvcc cross-compiles
    # for sm_100a without a GPU) instead of failing every test that loads it.  Not a fallback: the tests still run the library.
    from bayescard_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build(verbose=False)


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible and -m gpu was not asked for."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
