"""``Genred``: generic reductions with torch autograd -- API of ``pykeops.torch.Genred``
(``pykeops/pykeops/torch/generic/generic_red.py:386-695``), running on the sm_100a engine.

The backward pass follows the reference's protocol (``generic_red.py:152-259``): for every input
that needs a gradient, a new reduction string ``Grad_WithSavedForward(<formula>, <var>, <eta>, <res>)``
is handed to the engine, with the upstream gradient and the saved forward appended as two extra
variables.  Since each gradient is itself a differentiable ``Genred`` call, higher-order
derivatives work.
"""
from __future__ import annotations

import copy
from typing import List, Optional

import torch

from ..binder import LoadKeOpsB200
from ..common import (axis2cat, complete_aliases, get_optional_flags, get_sizes, get_type, postprocess,
                      preprocess)

_AD_UNSUPPORTED = ("Min_ArgMin_Reduction", "Min_Reduction", "Max_ArgMax_Reduction", "Max_Reduction",
                   "KMin_ArgKMin_Reduction", "KMin_Reduction")


def check_AD_supported(formula: str):
    for red in _AD_UNSUPPORTED:
        if formula.startswith(red):
            raise NotImplementedError(
                "As of today, KeOps does not support autodiff through the " + red + " reduction. "
                "You may want to extract the relevant integer indices with a '.argmin()', '.argmax()' or "
                "'.argKmin()' reduction before using PyTorch advanced indexing to create a fully-differentiable tensor.")


class GenredParameters:
    pass


def _resolve_backend(backend: str, args):
    """Only GPU backends exist here; ``"CPU"`` is refused (no fallback).  ``GPU_2D`` merely forces the split-column launch
    shape; the tuner may choose it on its own for ``auto`` / ``GPU_1D``.  Memory location as in the reference
    (``pykeops/pykeops/common/get_options.py:54-120``): all tensors on a CUDA device -> device mode; all on the host ->
    the ``*_host`` mode, in which the binder stages them on the GPU for the call and copies the result back (the reference's
    launcher does the same, ``keops_nvrtc.cpp:470-507,617-631``); a mixture is an error."""
    options = ("auto", "CPU", "GPU", "GPU_1D", "GPU_1D_device", "GPU_1D_host", "GPU_2D", "GPU_2D_device", "GPU_2D_host")
    if backend not in options:
        raise ValueError("Invalid backend. Should be one of ", options)
    if backend == "CPU":
        raise ValueError("[KeOps-B200] backend='CPU' is not available: this engine is CUDA (sm_100a) only.")
    if not all(isinstance(a, torch.Tensor) for a in args):
        raise TypeError("All variables should either be numpy arrays or torch tensors.")
    on_gpu = [a.is_cuda for a in args]
    if any(on_gpu) and not all(on_gpu):
        raise ValueError("At least two input variables have different memory locations (Cpu/Gpu).")
    host = not any(on_gpu)
    if backend.endswith("_device") and host or backend.endswith("_host") and not host:
        raise ValueError("[KeOps-B200] backend %r does not match the memory location of the inputs (%s)."
                         % (backend, "host" if host else "device"))
    if host and not torch.cuda.is_available():
        raise ValueError("[KeOps-B200] all inputs must be CUDA tensors (no CPU fallback): no CUDA device to stage host tensors on.")
    return 1, (1 if "2D" in backend else 0), (0 if host else 1)


def _set_device(device_id_request: int, args):
    """``set_device`` of the reference (``generic_red.py:46-67``): device tensors decide, host tensors go to the requested
    device (default: device 0)."""
    dev = args[0].device
    for a in args[1:]:
        if a.device != dev:
            raise ValueError("[KeOps] Input arrays must be all located on the same device.")
    if not dev.type == "cuda":
        return (0 if device_id_request in (-1, None) else int(device_id_request)), dev
    if device_id_request not in (-1, None) and dev.index != device_id_request:
        raise ValueError("[KeOps] Gpu device id of arrays is different from device id requested for computation.")
    return dev.index, dev


def _binder_for(params, args):
    flags = dict(params.optional_flags)
    flags["multVar_highdim"] = 1 if params.rec_multVar_highdim else 0
    tagCPUGPU, tag1D2D, tagHostDevice = _resolve_backend(params.backend, args)
    nbatchdims = max(len(a.shape) for a in args) - 2
    use_ranges = nbatchdims > 0 or params.ranges is not None
    device_id, device_args = _set_device(params.device_id_request, args)
    binder = LoadKeOpsB200(tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id, params.formula,
                           params.aliases, len(args), params.dtype, "torch", flags).import_module()
    return binder, device_args, nbatchdims


LAST_BINDER = None        # the binder of the most recent forward launch (introspection: bench.py, tests)


class GenredAutograd(torch.autograd.Function):
    """New-style autograd function (``forward`` + ``setup_context``) so that ``torch.func.jvp`` / ``torch.func.vmap`` work
    like with the reference (``generic_red.py:262-380``): ``backward`` hands ``Grad_WithSavedForward`` reductions to the
    engine, ``jvp`` hands ``Diff_WithSavedForward`` ones, ``vmap`` turns the mapped dimension into a batch dimension."""

    @staticmethod
    def forward(params, *args):
        global LAST_BINDER
        binder, device_args, nbatchdims = _binder_for(params, args)
        LAST_BINDER = binder
        cargs = tuple(a.detach() if a.is_contiguous() else a.detach().contiguous() for a in args)
        return binder.genred_pytorch(device_args, params.ranges, params.nx, params.ny, nbatchdims, params.out, *cargs)

    @staticmethod
    def setup_context(ctx, inputs, output):
        params, *args = inputs
        binder, _, _ = _binder_for(params, args)          # memoised: the binder built by forward
        ctx.params = params
        ctx.dimout, ctx.tagIJ = binder.dimout, binder.tagIJ
        ctx.save_for_backward(*args, output)
        ctx.save_for_forward(*args, output)

    @staticmethod
    def vmap(info, in_dims, params, *args):
        """The mapped dimension becomes a leading batch dimension of every argument (``generic_red.py:304-316``)."""
        moved = []
        for a, d in zip(args, in_dims[1:]):
            moved.append(a[None, ...] if d is None else a.movedim(d, 0).contiguous())
        p2 = copy.copy(params)
        p2.out = None
        return GenredAutograd.apply(p2, *moved), 0

    @staticmethod
    def jvp(ctx, _params_tangent, *tangents):
        """Forward-mode: one ``Diff_WithSavedForward`` reduction per input that carries a tangent
        (``generic_red.py:318-378``)."""
        params = ctx.params
        args = ctx.saved_tensors[:-1]
        result = ctx.saved_tensors[-1].detach()
        nargs = len(args)
        check_AD_supported(params.formula)
        resvar = "Var(%d,%d,%d)" % (nargs + 1, ctx.dimout, ctx.tagIJ)
        out = None
        for var_ind, (sig, t) in enumerate(zip(params.aliases, tangents)):
            if t is None:
                continue
            _, cat, dim, pos = get_type(sig, position_in_list=var_ind)
            var = "Var(%d,%d,%d)" % (pos, dim, cat)
            eta = "Var(%d,%d,%d)" % (nargs, dim, cat)
            params_d = copy.copy(params)
            params_d.formula = "Diff_WithSavedForward(%s,%s,%s,%s)" % (params.formula, var, eta, resvar)
            params_d.aliases = list(params.aliases) + [eta, resvar]
            params_d.out = None
            params_d.rec_multVar_highdim = None
            d = GenredAutograd.apply(params_d, *args, t.contiguous(), result)
            out = d if out is None else out + d
        return out if out is not None else torch.zeros_like(result)

    @staticmethod
    def backward(ctx, G):
        params = ctx.params
        args = ctx.saved_tensors[:-1]
        result = ctx.saved_tensors[-1].detach()
        nargs = len(args)
        check_AD_supported(params.formula)
        eta = "Var(%d,%d,%d)" % (nargs, ctx.dimout, ctx.tagIJ)
        resvar = "Var(%d,%d,%d)" % (nargs + 1, ctx.dimout, ctx.tagIJ)
        G = G.contiguous()
        grads: List[Optional[torch.Tensor]] = []
        for var_ind, (sig, arg) in enumerate(zip(params.aliases, args)):
            if not ctx.needs_input_grad[var_ind + 1]:
                grads.append(None)
                continue
            _, cat, dim, pos = get_type(sig, position_in_list=var_ind)
            var = "Var(%d,%d,%d)" % (pos, dim, cat)
            params_g = copy.copy(params)
            params_g.formula = "Grad_WithSavedForward(%s,%s,%s,%s)" % (params.formula, var, eta, resvar)
            params_g.aliases = list(params.aliases) + [eta, resvar]
            params_g.out = None
            params_g.rec_multVar_highdim = None
            grad = GenredAutograd.apply(params_g, *args, G, result)
            if cat == 2:
                # gradient w.r.t. a parameter: the engine reduced over j (output indexed by i), the
                # remaining sum over i is done here (``generic_red.py:216-235``)
                grad = (1.0 * grad).sum(-2)
                extra = tuple(i for i, (a, b) in enumerate(zip(arg.shape[:-1], grad.shape[:-1])) if a < b)
            else:
                extra = tuple(i for i, (a, b) in enumerate(zip(arg.shape[:-2], grad.shape[:-2])) if a < b)
            if extra:
                grad = (1.0 * grad).sum(extra, keepdim=True)
            grads.append(grad.reshape(arg.shape))
        return (None, *grads)


def genred_apply(params, *args):
    return GenredAutograd.apply(params, *args)


class Genred:
    """``Genred(formula, aliases, reduction_op='Sum', axis=0, ...)`` -- see the reference docstring
    (``generic_red.py:387-531``); all keyword arguments are accepted with the same meaning."""

    def __init__(self, formula, aliases, reduction_op="Sum", axis=0, dtype=None, opt_arg=None, formula2=None,
                 cuda_type=None, dtype_acc="auto", use_double_acc=False, sum_scheme="auto", enable_chunks=True,
                 rec_multVar_highdim=False, use_fast_math=True):
        self.reduction_op = reduction_op
        red_internal, formula2 = preprocess(reduction_op, formula2)
        self.optional_flags = get_optional_flags(red_internal, dtype_acc, use_double_acc, sum_scheme,
                                                 enable_chunks, use_fast_math)
        str_opt = ",%s" % opt_arg if opt_arg else ""
        str_f2 = ",%s" % formula2 if formula2 else ""
        self.formula = "%s_Reduction(%s%s,%d%s)" % (red_internal, formula, str_opt, axis2cat(axis), str_f2)
        self.aliases = complete_aliases(self.formula, list(aliases))
        self.nargs = len(self.aliases)
        self.axis = axis
        self.opt_arg = opt_arg
        self.rec_multVar_highdim = rec_multVar_highdim

    def __call__(self, *args, backend="auto", device_id=-1, ranges=None, out=None):
        if len(args) != self.nargs:
            raise ValueError("Invalid number of arguments in call to Genred (should be %d and got %d)."
                             % (self.nargs, len(args)))
        dtype = str(args[0].dtype).split(".")[1]
        nx, ny = get_sizes(self.aliases, *args)
        nout, nred = (nx, ny) if self.axis == 1 else (ny, nx)
        if "Arg" in self.reduction_op and nred > 1.6e7 and dtype in ("float32", "float"):
            # indices travel as floats in the accumulator (``generic_red.py:667-679``)
            raise ValueError("size of input array is too large for Arg type reduction with single precision. Use double precision.")
        params = GenredParameters()
        params.formula = self.formula
        params.aliases = self.aliases
        params.backend = backend
        params.dtype = dtype
        params.device_id_request = device_id
        params.ranges = ranges
        params.optional_flags = self.optional_flags
        params.rec_multVar_highdim = self.rec_multVar_highdim
        params.nx, params.ny = nx, ny
        params.out = out
        res = genred_apply(params, *args)
        return postprocess(res, self.reduction_op, self.opt_arg)


class RawReduction:
    """A reduction given by its full reduction string -- what the reference's backward hands to the engine
    (``Grad_WithSavedForward(<reduction>, <var>, <eta>, <res>)``, ``generic_red.py:190-214``) -- as a differentiable callable
    returning the raw ``(n, dim)`` result (no post-processing).  Used by ``keops_b200.distributed`` to run a gradient
    reduction sharded by the differentiated variable's own index."""

    def __init__(self, red_formula: str, aliases, optional_flags=None, rec_multVar_highdim=False):
        self.formula = red_formula
        self.aliases = list(aliases)
        self.nargs = len(self.aliases)
        self.optional_flags = dict(optional_flags) if optional_flags is not None else \
            get_optional_flags("Sum", "auto", False, "auto", True, True)
        self.rec_multVar_highdim = rec_multVar_highdim

    def __call__(self, *args, backend="auto", device_id=-1, ranges=None, out=None):
        if len(args) != self.nargs:
            raise ValueError("Invalid number of arguments (should be %d and got %d)." % (self.nargs, len(args)))
        params = GenredParameters()
        params.formula, params.aliases = self.formula, self.aliases
        params.backend = backend
        params.dtype = str(args[0].dtype).split(".")[1]
        params.device_id_request = device_id
        params.ranges = ranges
        params.optional_flags = self.optional_flags
        params.rec_multVar_highdim = self.rec_multVar_highdim
        params.nx, params.ny = get_sizes(self.aliases, *args)
        params.out = out
        return genred_apply(params, *args)


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
