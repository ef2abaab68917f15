import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (oracle/oracle.c).  Checker only - never the product."""
    import pyoracle
    pyoracle.build()
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference behind oracle/ref_shim.cpp, when oracle/_ref/libref.so exists."""
    import pyoracle
    try:
        r = pyoracle.Ref()
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref/libref.so not built (no /root/reference on this box)")
    r.set_threads(1)
    return r


def load_golden(name):
    return dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))


@pytest.fixture(scope="session")
def golden_names():
    import cases
    return list(cases.cases().keys())


@pytest.fixture(scope="session")
def thsp():
    """The CUDA library through its C ABI.  Fails (does not skip) when the .so is missing."""
    import arm_spmv_b200
    arm_spmv_b200.load()
    return arm_spmv_b200


@pytest.fixture(scope="session")
def cuda(thsp):
    import torch
    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)
