import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Make sure libagr_b200.so exists (nvcc cross-compiles without a GPU)."""
    from animatablegaussians_b200 import _build
    return _build.build_extension(verbose=False)


@pytest.fixture(autouse=True)
def _no_tf32():
    """fp32 parity tests compare against fp32 references: keep cuDNN / cuBLAS out of TF32."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
