import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# When the reference tree is mounted (build container) make `import laplace` resolve to it *before*
# laplace_b200 is imported, so that the backend derives from the real plug-in base classes and the
# drop-in tests can drive the unmodified `Laplace(...)` front end.  On the GPU box the reference is
# absent and laplace_b200 falls back to its host-side mirrors.  LPB_NO_REFERENCE=1 forces the mirrors.
if os.environ.get("LPB_NO_REFERENCE") != "1":
    from oracle import ref_shim

    ref_shim.install()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference tree mounted at /root/reference")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.pt")
    return torch.load(path, weights_only=False)


@pytest.fixture
def cpu_kernels(monkeypatch):
    """Install the torch-CPU emulation of the native kernels (tests/cpu_kernels.py) so that the host
    logic of laplace_b200 can be exercised without a GPU.  Test-only."""
    from tests import cpu_kernels as ck

    ck.install(monkeypatch)
    return ck
