import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: full-size CPU oracle checks (tens of seconds)")


def rel_l2(a, b):
    import torch
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float(torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="session")
def have_cuda():
    import torch
    return torch.cuda.is_available()
