import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o   # test infrastructure: the CPU restatement of the reference path
    o.build()
    return o


@pytest.fixture(scope="session")
def fb2():
    """The product package (host mirror + C ABI).  Building is the job of
    __graft_entry__.build(); tests only build when the .so is missing."""
    so = os.path.join(ROOT, "ferrum-infer-rs_b200", "libfb2_w4a16.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    return importlib.import_module("ferrum-infer-rs_b200")


@pytest.fixture(scope="session")
def cuda(fb2):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    torch.cuda.set_device(0)
    return torch
