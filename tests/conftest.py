import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pat(shape, k):
    """SURVEY.md §8c pattern: ((7c + 13i + 29j + 3k + 11b) mod 31)/31."""
    import torch

    b, c, h, w = shape
    B, C, I, J = torch.meshgrid(torch.arange(b), torch.arange(c), torch.arange(h), torch.arange(w), indexing="ij")
    return ((7 * C + 13 * I + 29 * J + 3 * k + 11 * B) % 31).float() / 31


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
