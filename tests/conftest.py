import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Build the oracle (always cheap) and, if missing, the CUDA library (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    if not os.path.exists(os.path.join(ROOT, "sw_reaxff_b200", "librxb200.so")):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "sw_reaxff_b200", "csrc"), "all"])
    yield


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
