import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")
    config.addinivalue_line("markers", "slow: multi-million-row cases")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a device: skip them (instead of failing) when there is none, so that a plain `pytest tests`
    on the CPU build host is green; and give every GPU test a time limit so that a hung kernel cannot eat the box."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    skip = pytest.mark.skip(reason="needs a CUDA device (run under gpurun with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            if not has_gpu:
                item.add_marker(skip)
            elif item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(600 if "slow" in item.keywords else 240))
