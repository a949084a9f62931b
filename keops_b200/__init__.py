"""keops_b200 -- a B200-native (sm_100a) implementation of the KeOps symbolic map-reduce hot path.

``out_i = Reduce_j F(x_i, y_j, p)`` for Sum / LogSumExp / Min-ArgMin / ArgKMin reductions and their
autodiff-generated gradient reductions, behind the ``pykeops.torch`` API (``keops_b200.torch``).
There is no CPU fallback: the CUDA runtime ``libkb200.so`` must be built (``__graft_entry__.build()``).
"""
__version__ = "0.1.0"

from .runtime import capi as _capi  # noqa: F401  (fails loudly if the native runtime is missing)
from .binder import LoadKeOpsB200, get_binder, register_into_pykeops  # noqa: F401
from .runtime.plan import clean_cache  # noqa: F401


def install_as_pykeops():
    """Make ``import pykeops`` / ``from pykeops.torch import LazyTensor, Genred, KernelSolve`` resolve to
    this package, so existing user code runs unchanged on the B200 engine."""
    import sys
    import types
    from . import numpy as _n
    from . import torch as _t
    from .numpy import cluster as _nc
    from .torch import cluster as _tc

    if "pykeops" in sys.modules and not getattr(sys.modules["pykeops"], "__kb200_shim__", False):
        raise RuntimeError("a real pykeops is already imported; use keops_b200.register_into_pykeops() instead")
    root = types.ModuleType("pykeops")
    root.__kb200_shim__ = True
    root.__version__ = "2.3+kb200." + __version__
    root.torch = _t
    root.numpy = _n
    root.clean_pykeops = lambda recompile_jit_binaries=True: (clean_cache(), LoadKeOpsB200.reset())[0]
    root.verbose = True
    root.default_device_id = 0
    root.set_verbose = lambda val: setattr(root, "verbose", val)
    root.set_build_folder = _set_build_folder
    root.get_build_folder = _get_build_folder
    root.check_health = _check_health
    root.config = types.SimpleNamespace(gpu_available=_cuda_available(), torch_found=True, numpy_found=True)
    root.test_torch_bindings = lambda: _smoke_install()
    root.test_numpy_bindings = lambda: _smoke_install(numpy=True)
    sys.modules["pykeops"] = root
    sys.modules["pykeops.torch"] = _t
    sys.modules["pykeops.numpy"] = _n
    sys.modules["pykeops.torch.cluster"] = _tc          # the very same module objects, not second copies found via __path__
    sys.modules["pykeops.numpy.cluster"] = _nc
    return root


def _cuda_available():
    import torch as th
    return bool(th.cuda.is_available())


def _set_build_folder(path=None):
    """``pykeops.set_build_folder`` (``pykeops/pykeops/__init__.py:83-93``): the folder that receives compiled kernels --
    here the cubin cache; ``None`` goes back to the default (in-tree) cache."""
    import os
    if path is None:
        os.environ.pop("KB200_CACHE_DIR", None)
        LoadKeOpsB200.reset()
    else:
        LoadKeOpsB200.reset(new_save_folder=path)


def _get_build_folder():
    from .runtime.plan import cache_dir
    return cache_dir()


def _check_health():
    """``pykeops.check_health`` (``pykeops/pykeops/__init__.py:64-80``): what this engine depends on, printed."""
    import torch as th
    from .runtime import capi
    print("keops_b200 %s: runtime library %s" % (__version__, capi.library_path() if hasattr(capi, "library_path") else "loaded"))
    print("  cubin cache: %s" % _get_build_folder())
    if th.cuda.is_available():
        for k in range(th.cuda.device_count()):
            cc = th.cuda.get_device_capability(k)
            print("  cuda:%d %s, compute capability %d.%d%s" % (k, th.cuda.get_device_name(k), cc[0], cc[1],
                                                                "" if cc == (10, 0) else "  (kernels are built for sm_100a only)"))
    else:
        print("  no CUDA device: every reduction will raise (there is no CPU backend)")


def _smoke_install(numpy=False):
    """Same known answer as ``pykeops/pykeops/torch/test_install.py:5-25`` / ``numpy/test_install.py:5-25``."""
    import torch as th
    if numpy:
        import numpy as np
        from .numpy import Genred as NpGenred
        x = np.arange(1, 10).reshape(-1, 3).astype("float32")
        y = np.arange(3, 9).reshape(-1, 3).astype("float32")
        out = NpGenred("SqNorm2(x-y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="Sum", axis=0)(x, y)
        ok = bool(np.allclose(out.reshape(-1), [63.0, 90.0]))
        print("pyKeOps-B200 with numpy bindings is working!" if ok else "pyKeOps-B200: outputs wrong values...")
        return ok
    from .torch import Genred
    x = th.arange(1, 10, dtype=th.float32).view(-1, 3).cuda()
    y = th.arange(3, 9, dtype=th.float32).view(-1, 3).cuda()
    out = Genred("SqNorm2(x-y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="Sum", axis=0)(x, y)
    ok = th.allclose(out.view(-1).cpu(), th.tensor([63.0, 90.0]))
    print("pyKeOps-B200 bindings are working!" if ok else "pyKeOps-B200: outputs wrong values...")
    return ok
