import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # built artefacts travel to the GPU box; in a fresh checkout build them once
    libs = [os.path.join(ROOT, "ospray_b200", "lib", n) for n in ("libospcuda.so", "libospray.so", "libospray_module_cuda.so")]
    libs.append(os.path.join(ROOT, "oracle", "liboracle.so"))
    if not all(os.path.exists(p) for p in libs):
        subprocess.check_call(["make", "-j8", "-C", ROOT], stdout=subprocess.DEVNULL)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def osp():
    """the cuda OSPRay device, initialised once per test session (GPU tests only)"""
    from ospray_b200 import capi
    capi.init()
    return capi
