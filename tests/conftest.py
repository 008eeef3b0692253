"""pytest configuration: `gpu` marker + shared paths.

`-m "not gpu"` runs here on CPU: oracle vs golden vectors, host logic, C-ABI symbol export.
`-m gpu` runs on a B200: the parity tests proper, all through the C-ABI library.
Nothing in the gpu tests reads /root/reference (it does not exist on the GPU box).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the package directory goes on sys.path the way a user of the reference's drivers would put it there (`from model import DGCNN`)
for p in (ROOT, os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "ahn3-dgcnn.pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
