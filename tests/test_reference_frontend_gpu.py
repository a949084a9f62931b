"""GPU: the drop-in claim, executed.  The reference's own front-end (``pykeops.torch``, staged unmodified into
``baseline/_ref`` by ``baseline/stage_ref.py``) and its own test files run on the B200 with this engine registered at the
reference's plugin seam ``keops_binder["nvrtc"]`` (``tests/reference_suite.py`` does the work in a separate process,
because ``test_torch_api_gpu.py`` installs ``keops_b200`` itself under the module name ``pykeops``).

Reference test bodies covered: ``pykeops/torch/test_install.py:5-25``, ``pykeops/test/test_torch.py:104-678`` (Genred float /
double / softmax / simple, LogSumExp + gradients, non-contiguity, heterogeneous aliases, KernelSolve ``test_invkernel``,
softmax, pickle, LazyTensor sum / logsumexp / min / min_chunked), ``test_lazytensor_gaussian_batch.py``,
``test_block_sparse_reduction.py:16-91``, ``test_lazytensor_grad.py``, ``test_torch_func.py``, ``test_vmap.py``, chunk /
final-chunk modes, ... -- the list is ``reference_suite.FILES``."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_test_files_pass_through_the_seam(cuda_device):
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "pykeops")):
        pytest.fail("baseline/_ref is not staged: run `python baseline/stage_ref.py` (or __graft_entry__.build()) in the "
                    "container where /root/reference is mounted -- the staged copy travels to the GPU box")
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    js = os.path.join(out, "reference_suite.json")
    log = os.path.join(out, "reference_suite.log")
    env = dict(os.environ)
    env.pop("KT_MODULE", None)
    with open(log, "w") as lf:
        proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "reference_suite.py"), "--json", js],
                              stdout=lf, stderr=subprocess.STDOUT, env=env, timeout=3000)
    assert os.path.exists(js), "reference suite did not report: see gpurun_out/reference_suite.log\n" + open(log).read()[-3000:]
    with open(js) as f:
        rep = json.loads(f.read())
    assert rep.get("failed") == [] and proc.returncode == 0, "failed: %s\n%s" % (rep.get("failed"), open(log).read()[-6000:])
    assert rep["passed"] >= 35, rep["passed"]
    assert rep["launches_through_LoadKeOpsB200"] >= 100, rep         # the CUDA path of this engine really did the work
