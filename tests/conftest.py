import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import vahana_b200 as vh  # noqa: E402

ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libvahana_oracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle as a vahana_b200 Backend (checker only)."""
    _build_oracle()
    b = vh.load_backend(ORACLE_LIB)
    assert b.name == "oracle-cpu"
    return b


@pytest.fixture(scope="session")
def cuda():
    """The product: the CUDA engine through its C-ABI.  No fallback."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    b = vh.default_backend()
    assert b.name.startswith("cuda")
    b.init(0)
    return b


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """Every semantic test runs against the oracle (CPU, pins the restatement to the reference's golden values)
    and, marked gpu, against the CUDA engine (same golden values through the C-ABI)."""
    return request.getfixturevalue(request.param)
