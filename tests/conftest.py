import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def tiny_sd():
    """Seeded TINY weights shared by oracle and product tests (same seed as tests/golden/make_golden.py)."""
    import torch
    from mr_blip_b200.dims import TINY, init_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    return init_state_dict(TINY, seed=1234, lora_b_std=0.02)
