"""Test configuration.  `-m "not gpu"` runs the oracle-vs-golden, host-logic and ABI tests on CPU;
`-m gpu` runs the parity tests proper (CUDA path through the C ABI vs the CPU oracle)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_binding():
    from oracle import binding
    binding.build()
    return binding
