import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _pin_global_generators():
    """torch's default seed is random per process (torch 2.11), and BoTorch-style code paths draw from the GLOBAL generator
    (initialize_q_batch's multinomial pick of the restarts, default seeds of strategies): pin it per test so that a test's
    outcome does not depend on the process it runs in or on the tests before it."""
    import numpy as np
    import torch

    torch.manual_seed(20250711)
    np.random.seed(20250711)
    yield
