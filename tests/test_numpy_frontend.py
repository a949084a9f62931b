"""CPU: the pykeops.numpy-compatible front-end (keops_b200/numpy) builds the same operation trees as the reference
front-end (the recordings of tests/golden/make_lazytensor_strings.py -- the reference's torch and numpy LazyTensors share
pykeops/pykeops/common/lazy_tensor.py), mirrors the torch front-end on data-carrying expressions, and refuses to compute
without a CUDA device (no CPU backend).  The numerical checks are in tests/test_numpy_api_gpu.py."""
import re

import numpy as np
import pytest
import torch

import keops_b200.numpy as kn
import keops_b200.torch as kt
from tests.test_lazytensor_strings import GEN, RECORDED, _same_tree


@pytest.fixture(scope="module")
def ours():
    return GEN.build(kn.LazyTensor)


@pytest.mark.parametrize("key", sorted(RECORDED))
def test_symbolic_trees_match_the_reference(key, ours):
    ref, got = RECORDED[key], ours[key]
    assert _same_tree(got["formula"], ref["formula"])
    if key.startswith("red:"):
        assert got["reduction_op"] == ref["reduction_op"] and got["axis"] == ref["axis"]
        assert str(got["opt_arg"]) == str(ref["opt_arg"])
        assert (got["formula2"] is None) == (ref["formula2"] is None)
    else:
        assert got["ndim"] == ref["ndim"]


def _relabel(f):
    seen = {}
    return re.sub(r"Var\((\d+),", lambda m: "Var(%d," % seen.setdefault(m.group(1), len(seen)), f)


def test_data_expressions_mirror_the_torch_front_end():
    rng = np.random.default_rng(0)
    x, y, b = (rng.random(s).astype("float32") for s in ((2, 5, 1, 3), (1, 1, 7, 3), (2, 1, 7, 2)))
    p = np.array([0.5], dtype="float32")

    def build(mod, conv):
        K = (-((mod.LazyTensor(conv(x)) - mod.LazyTensor(conv(y))) ** 2).sum(-1) * mod.Pm(conv(p)) / np.float32(2.0).item()).exp()
        r = (K * mod.LazyTensor(conv(b))).sum(2, call=False)
        return K, r

    Kn, rn = build(kn, lambda a: a)
    Kt, rt = build(kt, torch.from_numpy)
    assert type(Kn) is kn.LazyTensor and type(rn) is kn.LazyTensor and type(rn.callfun) is kn.Genred
    assert Kn.shape == Kt.shape == (2, 5, 7) and Kn.dtype == Kt.dtype == "float32"
    assert _relabel(Kn.formula) == _relabel(Kt.formula)
    assert rn.formula == rt.formula and rn.callfun.formula == rt.callfun.formula and rn.callfun.aliases == rt.callfun.aliases
    assert [tuple(v.shape) for v in rn.variables] == [tuple(v.shape) for v in rt.variables]
    assert Kn.T.shape == (2, 7, 5) and type(Kn.T) is kn.LazyTensor
    # numpy scalars behave like python numbers
    assert kn.LazyTensor(np.int64(3)).formula == "IntCst(3)"
    assert kn.LazyTensor(np.float32(0.5)).variables[0] == [0.5]


def test_constructors_mirror_the_torch_ones():
    a = dict(formula="Exp(-g*SqDist(x,y))*b", aliases=["x=Vi(3)", "y=Vj(3)", "b=Vj(2)", "g=Pm(1)"])
    gn, gt = kn.Genred(reduction_op="LogSumExp", axis=1, **a), kt.Genred(reduction_op="LogSumExp", axis=1, **a)
    assert (gn.formula, gn.aliases, gn.nargs, gn.optional_flags) == (gt.formula, gt.aliases, gt.nargs, gt.optional_flags)
    sn, st = kn.KernelSolve(a["formula"], a["aliases"], "b", axis=1), kt.KernelSolve(a["formula"], a["aliases"], "b", axis=1)
    assert (sn.formula, sn.aliases, sn.varinvpos) == (st.formula, st.aliases, st.varinvpos)
    k = kn.generic_argkmin("SqDist(x,y)", "a = Vi(5)", "x = Vi(3)", "y = Vj(3)")
    assert k.formula == "ArgKMin_Reduction(SqDist(x,y),5,0)" and k.opt_arg == 5
    assert kn.generic_sum("x|y", "a = Vj(1)", "x = Vi(3)", "y = Vj(3)").formula == "Sum_Reduction(x|y,1)"
    assert sorted(kn.__all__) == sorted(["Genred", "KernelSolve", "LazyTensor", "Vi", "Vj", "Pm", "generic_sum",
                                         "generic_logsumexp", "generic_argmin", "generic_argkmin"])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without a CUDA device")
def test_no_cpu_backend():
    x, y = np.zeros((4, 3), dtype="float32"), np.zeros((5, 3), dtype="float32")
    g = kn.Genred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="Sum", axis=1)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        g(x, y)
    with pytest.raises(ValueError, match="backend='CPU'"):
        g(x, y, backend="CPU")
    with pytest.raises(ValueError, match="Invalid number of arguments"):
        g(x)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        kn.Vi(x).sqdist(kn.Vj(y)).argKmin(2, dim=1)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        (-kn.Vi(x).sqdist(kn.Vj(x))).exp().solve(kn.Vi(x), alpha=0.1)


def test_pykeops_shim_exposes_numpy_and_cluster_modules(tmp_path):
    """``keops_b200.install_as_pykeops()``: the user's unchanged ``from pykeops.numpy import ...`` /
    ``from pykeops.torch.cluster import ...`` lines resolve to this package's module objects."""
    import sys
    import keops_b200
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "pykeops" or k.startswith("pykeops.")}
    try:
        root = keops_b200.install_as_pykeops()
        from pykeops.numpy import LazyTensor, Genred, Vi                      # noqa: F401
        from pykeops.numpy.cluster import grid_cluster as g_np
        from pykeops.torch.cluster import from_matrix as fm_t
        from pykeops.torch import LazyTensor as LT_t
        import keops_b200.numpy.cluster as nc
        import keops_b200.torch.cluster as tc
        assert LazyTensor is kn.LazyTensor and Genred is kn.Genred and LT_t is kt.LazyTensor
        assert g_np is nc.grid_cluster and fm_t is tc.from_matrix
        assert root.numpy is kn and callable(root.test_numpy_bindings)
        # housekeeping entry points user scripts call (pykeops/pykeops/__init__.py:20-23,47-104)
        import os
        import pykeops
        from keops_b200.runtime import plan
        pykeops.set_verbose(False)
        assert pykeops.verbose is False
        default = pykeops.get_build_folder()
        assert default == plan.cache_dir()
        prev = os.environ.get("KB200_CACHE_DIR")
        try:
            pykeops.set_build_folder(str(tmp_path / "kernels"))
            assert pykeops.get_build_folder() == str(tmp_path / "kernels") and os.path.isdir(pykeops.get_build_folder())
            pykeops.set_build_folder()
            assert pykeops.get_build_folder() == default or prev is not None
        finally:
            if prev is None:
                os.environ.pop("KB200_CACHE_DIR", None)
            else:
                os.environ["KB200_CACHE_DIR"] = prev
        pykeops.check_health()
        assert isinstance(pykeops.config.gpu_available, bool)
    finally:
        for k in [k for k in sys.modules if k == "pykeops" or k.startswith("pykeops.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_torch_front_end_memory_location_rules():
    """``backend`` / memory-location resolution of the torch front-end (pykeops/pykeops/common/get_options.py:54-120,
    torch/generic/generic_red.py:46-67): host tensors select the staged ``*_host`` mode on ``device_id`` (default 0) -- and
    raise where there is no CUDA device to stage them on; explicit ``*_device`` / ``*_host`` backends must match."""
    import torch
    from keops_b200.torch import genred
    x, y = torch.rand(4, 3), torch.rand(5, 3)
    assert genred._set_device(-1, (x, y)) == (0, x.device)
    assert genred._set_device(3, (x, y)) == (3, x.device)
    with pytest.raises(ValueError, match="backend='CPU' is not available"):
        genred._resolve_backend("CPU", (x, y))
    with pytest.raises(ValueError, match="Invalid backend"):
        genred._resolve_backend("TPU", (x, y))
    with pytest.raises(ValueError, match="does not match the memory location"):
        genred._resolve_backend("GPU_1D_device", (x, y))
    with pytest.raises(TypeError):
        genred._resolve_backend("auto", (x, y.numpy()))
    if not torch.cuda.is_available():
        with pytest.raises(ValueError, match="no CPU fallback"):
            genred._resolve_backend("auto", (x, y))
        with pytest.raises(ValueError, match="no CPU fallback"):
            kt.Genred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="Sum", axis=1)(x, y)
    else:
        assert genred._resolve_backend("auto", (x, y)) == (1, 0, 0)
        assert genred._resolve_backend("GPU_2D_host", (x, y)) == (1, 1, 0)
