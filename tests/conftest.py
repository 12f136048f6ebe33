import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: longer CPU test")


def _gpu_unavailable_reason():
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device"
    except Exception as exc:  # pragma: no cover
        return f"torch not importable: {exc!r}"
    from seigen_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        return f"{capi.LIB_PATH} not built"
    return None


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a GPU skips the gpu-marked tests instead of failing them (the product has no
    CPU fallback, so they cannot run).  With `-m gpu` on a box that should have a device nothing is skipped silently:
    a missing device / library there is an error of the run, reported by every test."""
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return                              # the GPU tier was asked for explicitly: let it fail loudly
    skip = pytest.mark.skip(reason=f"gpu test: {reason}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
