"""pytest configuration: markers, import paths, shared generators."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def fields():
    import fieldgen
    return fieldgen


_BIG = {}


def big_field(kind: str, size: int):
    """The 16384^2 / 8192^2 fields of the full-size tests, generated once per session (the smoothed one costs an
    FFT of 268 M points)."""
    import fieldgen
    key = (kind, size)
    if key not in _BIG:
        _BIG[key] = fieldgen.uniform(size, size, 0) if kind == "uniform" else fieldgen.smooth(size, size, 16.0, 0)
    return _BIG[key]
