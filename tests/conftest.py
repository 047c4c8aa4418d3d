import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_libs_built():
    """The shared libraries are git-ignored build products: on a fresh checkout build them once (nvcc cross-compiles
    sm_100a without a GPU; ~2-3 minutes) so the ABI / host-emulation / oracle tests do not fail for a missing file."""
    import subprocess
    so = os.path.join(ROOT, "zero_chain_b200", "libzkb200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "zero_chain_b200", "csrc"), "-j4", "-s"])
    oso = os.path.join(ROOT, "oracle", "libzkoracle.so")
    if not os.path.exists(oso):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    yield
