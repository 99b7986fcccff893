import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def _have_gpu():
    try:
        from chroma_lite_b200 import _lib
        return _lib.load().cb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; when they are selected without a
    # device they must fail loudly, not skip (no silent fallback).
    pass


@pytest.fixture(scope='session')
def gpu_ready():
    from chroma_lite_b200 import _lib
    assert _lib.load().cb_device_count() > 0, 'no CUDA device visible: GPU tests cannot run'
    _lib.init(0)
    return True
