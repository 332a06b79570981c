import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the C-ABI library is a build product (git-ignored): make sure it exists and is current
    # (nvcc cross-compiles sm_100a without a GPU; on the GPU box the prebuilt .so travels along)
    try:
        import torch
        from pychain_b200 import build as _b
        if not os.path.isfile(_b.LIB) or not torch.cuda.is_available():
            _b.build()  # no-op when the library is newer than its sources
    except Exception as e:  # leave the failure to the tests that need the library
        print("pychain_b200.build failed:", e)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
