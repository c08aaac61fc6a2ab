import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference checkout at /root/reference (authoring box only)")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    from oracle import ref_shim
    has_ref = ref_shim.available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)
        if "reference" in item.keywords and not has_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; built on demand so a fresh checkout can run the CPU suite."""
    from goal_force_b200 import build, capi
    build.build()
    return capi.load()


@pytest.fixture(scope="module")
def capi(lib):
    from goal_force_b200 import capi as c
    return c
