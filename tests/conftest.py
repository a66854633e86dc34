import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # fp32 parity: the two feature projections are library convs; cuDNN/cuBLAS would otherwise run them in TF32
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def isolated_rng():
    """Run a test on a forked copy of the global torch RNGs (CPU and CUDA): what it seeds or draws does not shift the random
    inputs of the tests that run after it (several older tests draw from the global stream)."""
    import torch
    with torch.random.fork_rng():
        yield
