import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def bf16_from_bits(bits):
    import torch

    return torch.from_numpy(bits.copy()).view(torch.bfloat16)


def f16_from_bits(bits, fp16):
    """int16 bit patterns -> torch.float16 (fp16 truthy) or torch.bfloat16 tensor."""
    import torch

    return torch.from_numpy(bits.copy()).view(torch.float16 if fp16 else torch.bfloat16)
