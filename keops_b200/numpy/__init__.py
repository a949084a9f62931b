"""``pykeops.numpy``-compatible front-end (``Genred``, ``KernelSolve``, ``LazyTensor``, ``Vi / Vj / Pm``, ``generic_*``)
over the sm_100a engine.

The reference's numpy binding (``pykeops/pykeops/numpy/generic/generic_red.py:52-68,188-330``,
``pykeops/pykeops/numpy/operations.py:47-61,179-290``, ``pykeops/pykeops/numpy/lazytensor/LazyTensor.py``) hands HOST
arrays to the launcher, which stages them on the GPU for the duration of the call (its ``GPU_*_host`` modes,
``keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:460-520``).  Here that staging is explicit: every call copies its
arrays to the selected CUDA device (``device_id``, default: the current device), runs the same kernels as
:mod:`keops_b200.torch` and copies the result back into a numpy array.  There is no CPU backend: without a CUDA device
(or with ``backend="CPU"``) every call raises.

Symbolic work (building formulas, transposition, ``K.shape`` ...) touches no device and works anywhere.
"""
from __future__ import annotations

import numpy as np
import torch

from ..common import get_type
from ..torch.genred import Genred as _TorchGenred
from ..torch.kernel_solve import KernelSolve as _TorchKernelSolve
from ..torch.lazytensor import LazyTensor as _TorchLazyTensor

__all__ = sorted(["Genred", "KernelSolve", "LazyTensor", "Vi", "Vj", "Pm", "generic_sum", "generic_logsumexp",
                  "generic_argmin", "generic_argkmin"])


def _device(device_id):
    if not torch.cuda.is_available():
        raise RuntimeError("[KeOps-B200] no CUDA device: this engine is CUDA (sm_100a) only, there is no CPU backend.")
    return torch.device("cuda", torch.cuda.current_device() if device_id in (-1, None) else int(device_id))


def _host_tensor(a):
    """numpy array / torch tensor -> torch tensor sharing the same memory (no copy for contiguous arrays)"""
    if isinstance(a, torch.Tensor):
        return a
    if not isinstance(a, np.ndarray):
        raise TypeError("[KeOps-B200] expected numpy arrays, got %s." % type(a))
    return torch.from_numpy(np.ascontiguousarray(a))


def _stage(args, backend, device_id):
    if backend == "CPU":
        raise ValueError("[KeOps-B200] backend='CPU' is not available: this engine is CUDA (sm_100a) only.")
    dev = _device(device_id)
    backend = backend[:-5] if backend.endswith("_host") else backend      # staging is what this function does
    return [_host_tensor(a).to(dev, non_blocking=True) for a in args], backend, dev


def _ranges_to(ranges, dev):
    return None if ranges is None else tuple(_host_tensor(np.asarray(r) if not isinstance(r, torch.Tensor) else r)
                                             .to(dev) for r in ranges)


def _back(res):
    if isinstance(res, tuple):
        return tuple(_back(r) for r in res)
    return res.detach().cpu().numpy()


class Genred:
    """``Genred(formula, aliases, reduction_op="Sum", axis=0, ...)(*arrays, backend="auto", device_id=-1, ranges=None,
    out=None)`` on numpy arrays (``pykeops/pykeops/numpy/generic/generic_red.py:52-68,188``)."""

    def __init__(self, formula, aliases, reduction_op="Sum", axis=0, dtype=None, opt_arg=None, formula2=None,
                 cuda_type=None, dtype_acc="auto", use_double_acc=False, sum_scheme="auto", enable_chunks=True,
                 rec_multVar_highdim=False, use_fast_math=True):
        self._impl = _TorchGenred(formula, aliases, reduction_op=reduction_op, axis=axis, dtype=dtype, opt_arg=opt_arg,
                                  formula2=formula2, cuda_type=cuda_type, dtype_acc=dtype_acc,
                                  use_double_acc=use_double_acc, sum_scheme=sum_scheme, enable_chunks=enable_chunks,
                                  rec_multVar_highdim=rec_multVar_highdim, use_fast_math=use_fast_math)
        for k in ("formula", "aliases", "nargs", "axis", "opt_arg", "reduction_op", "optional_flags"):
            setattr(self, k, getattr(self._impl, k))

    def __call__(self, *args, backend="auto", device_id=-1, ranges=None, out=None):
        if len(args) != self.nargs:
            raise ValueError("Invalid number of arguments in call to Genred (should be %d and got %d)."
                             % (self.nargs, len(args)))
        dargs, backend, dev = _stage(args, backend, device_id)
        res = _back(self._impl(*dargs, backend=backend, device_id=dev.index, ranges=_ranges_to(ranges, dev)))
        if out is not None:
            np.copyto(out, res)
            return out
        return res


class KernelSolve:
    """``KernelSolve(formula, aliases, varinvalias, axis=0, ...)(*arrays, alpha=1e-10, eps=1e-6)`` on numpy arrays
    (``pykeops/pykeops/numpy/operations.py:47-61,179-181``); the conjugate-gradient vectors stay on the device."""

    def __init__(self, formula, aliases, varinvalias, axis=0, dtype=None, opt_arg=None, dtype_acc="auto",
                 use_double_acc=False, sum_scheme="auto", enable_chunks=True, rec_multVar_highdim=None,
                 use_fast_math=True):
        self._impl = _TorchKernelSolve(formula, aliases, varinvalias, axis=axis, dtype_acc=dtype_acc,
                                       use_double_acc=use_double_acc, sum_scheme=sum_scheme,
                                       enable_chunks=enable_chunks, rec_multVar_highdim=rec_multVar_highdim,
                                       dtype=dtype, use_fast_math=use_fast_math)
        for k in ("formula", "aliases", "varinvpos", "axis", "optional_flags"):
            setattr(self, k, getattr(self._impl, k))

    def __call__(self, *args, backend="auto", device_id=-1, alpha=1e-10, eps=1e-6, ranges=None):
        dargs, backend, dev = _stage(args, backend, device_id)
        return _back(self._impl(*dargs, backend=backend, device_id=dev.index, alpha=alpha, eps=eps,
                                ranges=_ranges_to(ranges, dev)))


class LazyTensor(_TorchLazyTensor):
    """Symbolic tensors built from numpy arrays (``pykeops/pykeops/numpy/lazytensor/LazyTensor.py``): same operations
    and reductions as :class:`keops_b200.torch.LazyTensor`, reductions return numpy arrays."""
    _Genred = Genred
    _KernelSolve = KernelSolve

    @staticmethod
    def _ingest(x):
        if isinstance(x, np.generic):                   # numpy scalars behave like python numbers
            return x.item()
        return torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x

    @staticmethod
    def _view_as_complex(res):
        """interleaved (re, im) pairs -> complex array (``pykeops/pykeops/numpy/utils.py``, ``view_as_complex``)"""
        if isinstance(res, np.ndarray) and not np.iscomplexobj(res):
            res = np.ascontiguousarray(res)
            return res.view({"float32": "complex64", "float64": "complex128"}[str(res.dtype)])
        return res


def Vi(x_or_ind, dim=None):
    """``Vi(x)`` for an (M, D) array, or the symbolic ``Vi(ind, dim)``."""
    return LazyTensor((x_or_ind, dim, 0)) if isinstance(x_or_ind, int) else LazyTensor(x_or_ind, axis=0)


def Vj(x_or_ind, dim=None):
    return LazyTensor((x_or_ind, dim, 1)) if isinstance(x_or_ind, int) else LazyTensor(x_or_ind, axis=1)


def Pm(x_or_ind, dim=None):
    if isinstance(x_or_ind, int) and dim is not None:
        return LazyTensor((x_or_ind, dim, 2))
    return LazyTensor(x_or_ind, axis=2) if isinstance(x_or_ind, np.ndarray) else LazyTensor(x_or_ind)


def generic_sum(formula, output, *aliases, **kwargs):
    _, cat, _, _ = get_type(output)
    return Genred(formula, aliases, reduction_op="Sum", axis=1 - cat if cat in (0, 1) else 0, **kwargs)


def generic_logsumexp(formula, output, *aliases, **kwargs):
    _, cat, _, _ = get_type(output)
    return Genred(formula, aliases, reduction_op="LogSumExp", axis=1 - cat, **kwargs)


def generic_argkmin(formula, output, *aliases, **kwargs):
    _, cat, k, _ = get_type(output)
    return Genred(formula, aliases, reduction_op="ArgKMin", axis=1 - cat, opt_arg=k, **kwargs)


def generic_argmin(formula, output, *aliases, **kwargs):
    _, cat, _, _ = get_type(output)
    return Genred(formula, aliases, reduction_op="ArgMin", axis=1 - cat, **kwargs)
