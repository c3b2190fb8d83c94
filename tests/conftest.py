import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a box without CUDA or without the built extension, so that a plain
    `pytest tests` on a CPU box shows real CPU regressions only.  On a GPU box the extension MUST load: a missing
    library is a failure there (the product has no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run with `pytest -m gpu` on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "small_e2w.pt"), weights_only=False)
