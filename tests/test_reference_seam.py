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


def _run_reference(script, tmp_path, tag, timeout, cache=None):
    """Run ``script`` with the real reference importable; returns the JSON payload printed after ``tag``.  An environment
    in which the reference itself cannot run (no compiler, read-only cache ...) skips; a failed assertion fails."""
    if cache is None:
        cache = os.environ.get("KEOPS_CACHE_FOLDER", os.path.join(ROOT, "oracle", "_ref", "keops_cache"))   # git-ignored, persistent
    os.makedirs(cache, exist_ok=True)
    env = dict(os.environ, CXX="/usr/bin/g++", KEOPS_CACHE_FOLDER=cache,
               PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=os.pathsep.join((REF + "/keopscore", REF + "/pykeops", ROOT)))
    r = subprocess.run([sys.executable, "-c", script], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        if "AssertionError" in r.stderr or "keops_b200" in r.stderr.split("Traceback")[-1]:
            pytest.fail(r.stdout[-2000:] + r.stderr[-2000:])
        pytest.skip("the reference could not run in this environment: " + (r.stderr.strip().splitlines() or ["?"])[-1][:200])
    return json.loads([l for l in r.stdout.splitlines() if l.startswith(tag)][-1][len(tag):])

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
    recs = _run_reference(SCRIPT, tmp_path, "SEAM", 900)
    assert len(recs) == 6
    kinds = [rec["formula"].split("_Reduction")[0] for rec in recs]
    assert kinds == ["Sum", "ArgKMin", "Max_SumShiftExp", "Min_ArgMin", "Sum", "Sum"]
    assert recs[4]["flags"]["sum_scheme"] == "kahan_scheme" and recs[4]["flags"]["use_fast_math"] == 0
    for rec in recs:
        assert rec["ours"] == rec["theirs"], rec
        assert rec["cpu_refused"]                            # the reference's CPU argument list is refused: no CPU path here


# ---------------------------------------------------------------------------------------------------------------------
# call-time contract: what each front-end hands to ``binder.genred_pytorch`` for the same user calls
# ---------------------------------------------------------------------------------------------------------------------
USER_CALLS = r'''
def user_calls(torch, Genred, LazyTensor):
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.rand(*s, generator=g)
    x, y, b = r(10, 3), r(12, 3), r(12, 2)
    conv = Genred("Exp(-SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)"], reduction_op="Sum", axis=1)
    ranges = (torch.tensor([[0, 4], [4, 10]]), torch.tensor([1, 2]), torch.tensor([[0, 5], [5, 12]]),
              torch.tensor([[0, 5], [5, 12]]), torch.tensor([1, 2]), torch.tensor([[0, 4], [4, 10]]))
    xb, yb, bb = r(4, 1, 7, 1, 3), r(1, 5, 1, 9, 3), r(4, 5, 1, 9, 2)
    return [
        ("plain", lambda: conv(x, y, b)),
        ("out", lambda: conv(x, y, b, out=torch.empty(10, 2))),
        ("ranges", lambda: conv(x, y, b, ranges=ranges)),
        ("axis0", lambda: Genred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="ArgKMin", axis=0, opt_arg=2)(x, y)),
        ("batch", lambda: ((-((LazyTensor(xb) - LazyTensor(yb)) ** 2).sum(-1)).exp() * LazyTensor(bb)).sum(3)),
        ("batch_genred", lambda: conv(r(6, 10, 3), r(1, 12, 3), r(6, 12, 2))),
        ("param", lambda: Genred("Exp(-p*SqDist(x,y))", ["x=Vi(3)", "y=Vj(3)", "p=Pm(1)"], axis=1)(x, y, torch.tensor([0.5]))),
    ]

def describe(device_args, ranges, nx, ny, nbatchdims, out, *args):
    return dict(ranges=None if ranges is None else [list(t.shape) for t in ranges], nx=int(nx), ny=int(ny),
                nbatchdims=int(nbatchdims), out=None if out is None else list(out.shape),
                args=[list(a.shape) for a in args], contiguous=all(a.is_contiguous() for a in args))
'''

REF_CALLS = USER_CALLS + r'''
import json, torch
from pykeops.common import keops_io
class Stop(Exception):
    pass
seen = []
class Spy:
    def __init__(self, *a):
        pass
    def import_module(self):
        return self
    def genred_pytorch(self, *a):
        seen.append(describe(*a))
        raise Stop
keops_io.keops_binder["cpp"] = Spy
from pykeops.torch import Genred, LazyTensor
out = {}
for name, f in user_calls(torch, Genred, LazyTensor):
    try:
        f()
    except Stop:
        out[name] = seen[-1]
print("CALLS" + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pykeops")), reason="reference checkout not mounted")
def test_both_front_ends_hand_the_same_call_to_the_binder(tmp_path, monkeypatch):
    theirs = _run_reference(REF_CALLS, tmp_path, "CALLS", 900)
    assert len(theirs) == 7

    # the same user calls through this package's front-end, stopped at the binder in the same way (the device checks
    # that refuse CPU tensors are bypassed: nothing is computed)
    import torch
    import keops_b200.torch as kt
    from keops_b200.binder import LoadKeOpsB200
    from keops_b200.torch import genred as G
    ns = {}
    exec(USER_CALLS, ns)

    class Stop(Exception):
        pass
    seen = []

    def record(self, *a):
        seen.append(ns["describe"](*a))
        raise Stop
    monkeypatch.setattr(G, "_resolve_backend", lambda backend, args: (1, 0, 1))
    monkeypatch.setattr(G, "_set_device", lambda req, args: (0, torch.device("cpu")))
    monkeypatch.setattr(LoadKeOpsB200, "genred_pytorch", record)
    ours = {}
    for name, f in ns["user_calls"](torch, kt.Genred, kt.LazyTensor):
        with pytest.raises(Stop):
            f()
        ours[name] = seen[-1]
    for name in theirs:
        assert ours[name] == theirs[name], (name, ours[name], theirs[name])


# ---------------------------------------------------------------------------------------------------------------------
# autograd glue: which reductions each front-end asks the binder for during backward
# ---------------------------------------------------------------------------------------------------------------------
GRAD_CALLS = r'''
def grad_calls(torch, Genred, LazyTensor):
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.rand(*s, generator=g).requires_grad_(True)
    x, y, b, p = r(10, 3), r(12, 3), r(12, 2), torch.tensor([0.5], requires_grad=True)
    def conv():
        out = Genred("Exp(-p*SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)", "p=Pm(1)"], reduction_op="Sum", axis=1)(x, y, b, p)
        torch.autograd.grad(out.sum(), [x, y, b, p])
    def lse():
        out = (-((LazyTensor(x[:, None]) - LazyTensor(y[None])) ** 2).sum(-1)).logsumexp(dim=1, weight=LazyTensor(b[None]))
        torch.autograd.grad(out.sum(), [x, b])
    def second_order():
        out = ((-((LazyTensor(x[:, None]) - LazyTensor(y[None])) ** 2).sum(-1)).exp() * LazyTensor(b[None])).sum(dim=1)
        gx, = torch.autograd.grad((out ** 2).sum(), [x], create_graph=True)
        torch.autograd.grad((gx ** 2).sum(), [b])
    return [("conv", conv), ("lse", lse), ("second_order", second_order)]
'''

REF_GRADS = GRAD_CALLS + r'''
import json, torch
from pykeops.common import keops_io
real_cpp = keops_io.keops_binder["cpp"]
seen = []
def spy(*a):
    seen.append([a[5], list(a[6]), a[7], a[8], a[10]["sum_scheme"]])
    return real_cpp(*a)
keops_io.keops_binder["cpp"] = spy
from pykeops.torch import Genred, LazyTensor
out = {}
for name, f in grad_calls(torch, Genred, LazyTensor):
    start = len(seen)
    f()
    out[name] = seen[start:]
print("GRADS" + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pykeops")), reason="reference checkout not mounted")
def test_backward_asks_the_binder_for_the_same_reductions(tmp_path, monkeypatch):
    theirs = _run_reference(REF_GRADS, tmp_path, "GRADS", 1500)

    # ours: the binder is replaced by a recorder whose "kernel" returns zeros of the right shape, so that autograd runs
    import torch
    import keops_b200.torch as kt
    from keops_b200.binder import LoadKeOpsB200
    from keops_b200.torch import genred as G
    ns = {}
    exec(GRAD_CALLS, ns)
    seen = []

    def recording_binder(*a):
        seen.append([a[5], list(a[6]), a[7], a[8], a[10]["sum_scheme"]])
        return LoadKeOpsB200(*a)

    def fake_kernel(self, device_args, ranges, nx, ny, nbatchdims, out, *args):
        return torch.zeros(nx if self.tagIJ == 0 else ny, self.dimout, dtype=args[0].dtype)
    monkeypatch.setattr(G, "LoadKeOpsB200", recording_binder)
    monkeypatch.setattr(G, "_resolve_backend", lambda backend, args: (1, 0, 1))
    monkeypatch.setattr(G, "_set_device", lambda req, args: (0, torch.device("cpu")))
    monkeypatch.setattr(LoadKeOpsB200, "genred_pytorch", fake_kernel)
    from keops_b200.formulas.parser import parse_reduction
    for name, f in ns["grad_calls"](torch, kt.Genred, kt.LazyTensor):
        start = len(seen)
        f()
        mine = seen[start:]
        # setup_context looks the (memoised) binder up a second time: drop immediate repeats
        mine = [c for k, c in enumerate(mine) if k == 0 or c != mine[k - 1]]
        ref = [c for k, c in enumerate(theirs[name]) if k == 0 or c != theirs[name][k - 1]]
        assert len(mine) == len(ref), (name, [c[0] for c in mine], [c[0] for c in ref])
        for a, b in zip(mine, ref):
            assert a[1:] == b[1:], (name, a, b)                              # aliases, nargs, dtype, sum scheme
            assert parse_reduction(a[0], a[1]) is parse_reduction(b[0], b[1]) or a[0] == b[0], (name, a[0], b[0])


KSOLVE_CALLS = r'''
def ksolve_call(torch, KernelSolve):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(15, 3, generator=g).requires_grad_(True)
    b = torch.rand(15, 2, generator=g).requires_grad_(True)
    p = torch.tensor([0.5], requires_grad=True)
    Kinv = KernelSolve("Exp(-p*SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)", "p=Pm(1)"], "b", axis=1)
    c = Kinv(x, x, b, p, alpha=0.1)
    torch.autograd.grad(c.sum(), [x, b, p])
'''

REF_KSOLVE = KSOLVE_CALLS + r'''
import json, torch
from pykeops.common import keops_io
real_cpp = keops_io.keops_binder["cpp"]
seen = []
def spy(*a):
    seen.append([a[5], list(a[6]), a[7], a[8], a[10]["sum_scheme"]])
    return real_cpp(*a)
keops_io.keops_binder["cpp"] = spy
from pykeops.torch import KernelSolve
ksolve_call(torch, KernelSolve)
print("KSOLVE" + json.dumps(seen))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pykeops")), reason="reference checkout not mounted")
def test_kernel_solve_backward_asks_for_the_same_reductions(tmp_path, monkeypatch):
    """``KernelSolveAutograd`` (pykeops/pykeops/torch/operations.py:29-180): forward solve, a second solve for the
    upstream gradient, then one ``Grad_WithSavedForward`` reduction per differentiated variable."""
    theirs = _run_reference(REF_KSOLVE, tmp_path, "KSOLVE", 1500)
    import torch
    import keops_b200.torch as kt
    from keops_b200.binder import LoadKeOpsB200
    from keops_b200.torch import genred as G, kernel_solve as KS
    ns = {}
    exec(KSOLVE_CALLS, ns)
    seen = []

    def recording_binder(*a):
        seen.append([a[5], list(a[6]), a[7], a[8], a[10]["sum_scheme"]])
        return LoadKeOpsB200(*a)

    def fake_kernel(self, device_args, ranges, nx, ny, nbatchdims, out, *args):   # K = 0: the solver sees alpha * Id
        return torch.zeros(nx if self.tagIJ == 0 else ny, self.dimout, dtype=args[0].dtype)
    for mod in (G, KS):
        monkeypatch.setattr(mod, "LoadKeOpsB200", recording_binder)
        monkeypatch.setattr(mod, "_resolve_backend", lambda backend, args: (1, 0, 1))
        monkeypatch.setattr(mod, "_set_device", lambda req, args: (0, torch.device("cpu")))
    monkeypatch.setattr(LoadKeOpsB200, "genred_pytorch", fake_kernel)
    ns["ksolve_call"](torch, kt.KernelSolve)
    from keops_b200.formulas.parser import parse_reduction
    dedup = lambda calls: [c for k, c in enumerate(calls) if k == 0 or c != calls[k - 1]]
    mine, ref = dedup(seen), dedup(theirs)
    key = lambda c: (c[0], tuple(c[1]))
    assert sorted(map(key, mine)) == sorted(map(key, ref)) or len(mine) == len(ref), ([c[0] for c in mine], [c[0] for c in ref])
    for a, b in zip(sorted(mine, key=key), sorted(ref, key=key)):
        assert a[1:] == b[1:], (a, b)
        assert a[0] == b[0] or parse_reduction(a[0], a[1]) is parse_reduction(b[0], b[1]), (a[0], b[0])


FUNC_CALLS = r'''
def func_calls(torch, Genred):
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.rand(*s, generator=g)
    x, y, b = r(10, 3), r(12, 3), r(12, 2)
    conv = Genred("Exp(-SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)"], reduction_op="Sum", axis=1)
    lse = Genred("-SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="LogSumExp", axis=1)
    def jvp():
        torch.func.jvp(lambda x_, b_: conv(x_, y, b_), (x, b), (r(10, 3), r(12, 2)))
    def jvp_lse():
        torch.func.jvp(lambda x_: lse(x_, y), (x,), (r(10, 3),))
    def vmap():
        torch.func.vmap(lambda x_, b_: conv(x_, y, b_), in_dims=(0, 0))(r(4, 10, 3), r(4, 12, 2))
    def vmap_grad():
        torch.func.vmap(torch.func.grad(lambda x_: conv(x_, y, b).sum()))(r(4, 10, 3))
    return [("jvp", jvp), ("jvp_lse", jvp_lse), ("vmap", vmap), ("vmap_grad", vmap_grad)]
'''

REF_FUNC = FUNC_CALLS + r'''
import json, torch
from pykeops.common import keops_io
real_cpp = keops_io.keops_binder["cpp"]
seen = []
class Spy:
    def __init__(self, *a):
        self.a, self.real = a, real_cpp(*a)
    def import_module(self):
        self.real = self.real.import_module()
        return self
    def __getattr__(self, k):
        return getattr(self.real, k)
    def genred_pytorch(self, device_args, ranges, nx, ny, nbatchdims, out, *args):
        seen.append([self.a[5], list(self.a[6]), int(nx), int(ny), int(nbatchdims), [list(t.shape) for t in args]])
        return self.real.genred_pytorch(device_args, ranges, nx, ny, nbatchdims, out, *args)
keops_io.keops_binder["cpp"] = Spy
from pykeops.torch import Genred
out = {}
for name, f in func_calls(torch, Genred):
    start = len(seen)
    f()
    out[name] = seen[start:]
print("FUNC" + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pykeops")), reason="reference checkout not mounted")
def test_torch_func_transforms_launch_the_same_reductions(tmp_path, monkeypatch):
    """``torch.func.jvp / vmap / grad`` through ``GenredAutograd`` (pykeops/pykeops/torch/generic/generic_red.py:262-380):
    every kernel launch -- reduction string, aliases, nx, ny, number of batch dimensions, argument shapes -- is the one the
    real reference makes for the same transform."""
    theirs = _run_reference(REF_FUNC, tmp_path, "FUNC", 1500)
    import torch
    import keops_b200.torch as kt
    from keops_b200.binder import LoadKeOpsB200
    from keops_b200.torch import genred as G
    from keops_b200.formulas.parser import parse_reduction
    ns = {}
    exec(FUNC_CALLS, ns)
    seen = []

    def fake_kernel(self, device_args, ranges, nx, ny, nbatchdims, out, *args):
        p = self.params
        seen.append([p.red_formula_string, list(p.aliases_old), int(nx), int(ny), int(nbatchdims), [list(t.shape) for t in args]])
        batch = torch.broadcast_shapes(*[tuple(a.shape[:nbatchdims]) for a in args]) if nbatchdims else ()
        return torch.zeros(*batch, nx if self.tagIJ == 0 else ny, self.dimout, dtype=args[0].dtype)
    monkeypatch.setattr(G, "_resolve_backend", lambda backend, args: (1, 0, 1))
    monkeypatch.setattr(G, "_set_device", lambda req, args: (0, torch.device("cpu")))
    monkeypatch.setattr(LoadKeOpsB200, "genred_pytorch", fake_kernel)
    for name, f in ns["func_calls"](torch, kt.Genred):
        start = len(seen)
        f()
        mine, ref = seen[start:], theirs[name]
        assert len(mine) == len(ref), (name, [c[0] for c in mine], [c[0] for c in ref])
        for a, b in zip(mine, ref):
            assert a[2:] == b[2:], (name, a, b)                              # nx, ny, nbatchdims, argument shapes
            assert a[0] == b[0] or parse_reduction(a[0], a[1]) is parse_reduction(b[0], b[1]), (name, a[0], b[0])


HOUSEKEEPING = r'''
import json, os, sys, tempfile
import pykeops
from pykeops.common import keops_io
from keops_b200 import binder as kb
kb.register_into_pykeops()
fl = dict(dtype_acc="auto", sum_scheme="auto", use_fast_math=1, enable_chunks=1, multVar_highdim=0)
mk = lambda: kb.LoadKeOpsB200(1, 0, 1, False, -1, "Sum_Reduction(SqDist(x,y),0)", ["x=Vi(0,3)", "y=Vj(1,3)"], 2, "float32", "torch", fl)
b1 = mk()
assert mk() is b1                                       # memoised per argument tuple
b1.reset()                                              # the per-object form (pykeops/__init__.py:57-59 on a binder object)
pykeops.clean_pykeops(recompile_jit_binaries=False)     # calls keops_binder[key].reset() on every entry, ours included
b2 = mk()
assert b2 is not b1
folder = tempfile.mkdtemp(prefix="kb200_build_folder_")
pykeops.set_build_folder(folder)                        # calls keops_binder[key].reset(new_save_folder=...)
from keops_b200.runtime import plan
assert os.path.realpath(plan.cache_dir()) == os.path.realpath(folder), plan.cache_dir()
assert mk() is not b2
print("HOUSE" + json.dumps({"ok": True, "build_folder": pykeops.get_build_folder()}))
'''


def test_clean_pykeops_and_set_build_folder_reach_the_registered_binder(tmp_path):
    """``pykeops.clean_pykeops()`` and ``pykeops.set_build_folder()`` call ``.reset(...)`` on every ``keops_binder`` entry
    (pykeops/pykeops/__init__.py:47-61,83-93): with this engine registered that entry is the class itself."""
    # its own throw-away reference cache: clean_pykeops() empties the folder it is pointed at
    rec = _run_reference(HOUSEKEEPING, tmp_path, "HOUSE", 900, cache=str(tmp_path / "keops_cache"))
    assert rec["ok"]
