import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "libngscov_oracle.so")
    src = os.path.join(ROOT, "oracle", "ngscov_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return so


@pytest.fixture(scope="session")
def oracle():
    _ensure_oracle()
    from tests import oracle_py
    return oracle_py


@pytest.fixture(scope="session")
def ref_bin():
    """Paths of the unmodified reference binaries (oracle/_ref), or None when they were not built."""
    d = os.path.join(ROOT, "oracle", "_ref")
    tools = {k: os.path.join(d, k) for k in ("bam2depth", "bam2wig", "sam2bam")}
    if all(os.path.exists(p) for p in tools.values()):
        return tools
    return None


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
