"""The drop-in seam against the REAL reference front-end (CPU; skipped where the reference checkout is not mounted, i.e. on
the GPU box): after ``register_into_pykeops()`` the unmodified ``pykeops.torch.Genred / LazyTensor / KernelSolve`` find
``LoadKeOpsB200`` under ``keops_binder["nvrtc"]`` (pykeops/pykeops/common/keops_io/__init__.py:6-13), and the exact
argument lists they hand to a binder (generic_red.py:89-103, torch/operations.py:45-57) -- captured with a recording
stand-in under ``keops_binder["cpp"]``, since CPU tensors are routed there -- construct our binder, whose ``dimout`` /
``tagIJ`` / ``params`` agree with what the reference's own C++ binder (``LoadKeOps_cpp``) computes for the same list.
Runs in a subprocess so that the real ``pykeops`` never enters this test process."""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, torch
import pykeops
from pykeops.common import keops_io
from keops_b200.binder import register_into_pykeops, LoadKeOpsB200
real_cpp = keops_io.keops_binder["cpp"]
register_into_pykeops()
assert keops_io.keops_binder["nvrtc"] is LoadKeOpsB200
calls = []
class Stop(Exception):
    pass
def spy(*args):
    calls.append(args)
    raise Stop
keops_io.keops_binder["cpp"] = spy
from pykeops.torch import Genred, LazyTensor, KernelSolve
x, y, b = torch.rand(10, 3), torch.rand(12, 3), torch.rand(12, 2)
al = ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)"]
D = lambda: ((LazyTensor(x[:, None]) - LazyTensor(y[None])) ** 2).sum(-1)
for f in (lambda: Genred("Exp(-SqDist(x,y))*b", al, reduction_op="Sum", axis=1)(x, y, b),
          lambda: D().argKmin(3, dim=1),
          lambda: (-D()).logsumexp(dim=1),
          lambda: D().min_argmin(dim=0),
          lambda: ((-D()).exp() * LazyTensor(b[None])).sum(dim=1, sum_scheme="kahan_scheme", use_fast_math=False),
          lambda: KernelSolve("Exp(-SqDist(x,y))*b", al, "b", axis=1)(x, x, torch.rand(10, 2), alpha=0.1)):
    try:
        f()
    except Stop:
        pass
keops_io.keops_binder["cpp"] = real_cpp
out = []
for args in calls:
    assert args[0] == 0 and args[2] == 0                     # CPU tensors: tagCPUGPU = tagHostDevice = 0
    theirs = real_cpp(*args).import_module()                  # the reference's own binder, CPU code path
    gpu_args = (1, args[1], 1) + tuple(args[3:])              # what the same call passes for CUDA tensors
    ours = LoadKeOpsB200(*gpu_args).import_module()
    rec = dict(formula=args[5], flags=args[10], use_ranges=args[3])
    for who, bnd in (("ours", ours), ("theirs", theirs)):
        p = bnd.params
        rec[who] = dict(dimout=int(bnd.dimout), tagIJ=int(bnd.tagIJ), tagI=int(p.tagI), dim=int(p.dim), dimred=int(p.dimred),
                        dimy=int(p.dimy), indsi=[int(v) for v in p.indsi], indsj=[int(v) for v in p.indsj],
                        indsp=[int(v) for v in p.indsp], dimsx=[int(v) for v in p.dimsx], dimsy=[int(v) for v in p.dimsy],
                        dimsp=[int(v) for v in p.dimsp])
    try:
        LoadKeOpsB200(*args).import_module()
        rec["cpu_refused"] = False
    except ValueError:
        rec["cpu_refused"] = True
    out.append(rec)
print("SEAM" + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pykeops")), reason="reference checkout not mounted")
def test_real_reference_front_end_reaches_our_binder(tmp_path):
    env = dict(os.environ, CXX="/usr/bin/g++", KEOPS_CACHE_FOLDER=os.environ.get("KEOPS_CACHE_FOLDER", "/tmp/keops_cache"),
               PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=os.pathsep.join((REF + "/keopscore", REF + "/pykeops", ROOT)))
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    recs = json.loads([l for l in r.stdout.splitlines() if l.startswith("SEAM")][-1][4:])
    assert len(recs) == 6
    kinds = [rec["formula"].split("_Reduction")[0] for rec in recs]
    assert kinds == ["Sum", "ArgKMin", "Max_SumShiftExp", "Min_ArgMin", "Sum", "Sum"]
    assert recs[4]["flags"]["sum_scheme"] == "kahan_scheme" and recs[4]["flags"]["use_fast_math"] == 0
    for rec in recs:
        assert rec["ours"] == rec["theirs"], rec
        assert rec["cpu_refused"]                            # the reference's CPU argument list is refused: no CPU path here
