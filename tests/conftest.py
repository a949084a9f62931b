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


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Print (and save under gpurun_out/) the ACHIEVED element-wise parity errors, not only pass / fail."""
    try:
        from tests import util
    except Exception:
        return
    if not util.ACHIEVED:
        return
    import json
    worst = sorted(util.ACHIEVED.items(), key=lambda kv: -kv[1])
    terminalreporter.write_line("achieved parity (max |got - ref| / condition scale) -- %d checks, worst 12:" % len(worst))
    for what, v in worst[:12]:
        terminalreporter.write_line("  %.3e  %s" % (v, what))
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        tag = "gpu" if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "") else "cpu"
        with open(os.path.join(out, "parity_achieved_%s.json" % tag), "w") as f:
            json.dump(dict(worst), f, indent=0)
    except OSError:
        pass
