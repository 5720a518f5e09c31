import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device skips the gpu-marked tests instead of failing in Engine.__init__.
    On a GPU box nothing is skipped: a missing libd3dp_b200.so must FAIL there (there is no fallback to hide behind)."""
    import torch
    if not torch.cuda.is_available():
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def engine27():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from d3dp_b200.engine import Engine
    return Engine(frames=27)
