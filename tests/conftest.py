import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass on a box without a GPU: they are skipped only when the
    user did not ask for them (``-m gpu`` on a GPU-less machine fails loudly in the fixture)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    selected_gpu = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    if selected_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if os.environ.get("KB200_TEST_DEVICE") == "cpu":
        # only for running the front-end test BODIES unchanged against the real reference (KT_MODULE=pykeops.torch /
        # KN_MODULE=pykeops.numpy, CPU backend) to pin their expectations; keops_b200 itself raises on CPU tensors
        return torch.device("cpu")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda:0")
