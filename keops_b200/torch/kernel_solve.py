"""``KernelSolve``: conjugate-gradient solver around a generic reduction -- API of
``pykeops.torch.KernelSolve`` (``pykeops/pykeops/torch/operations.py:168-406``).

Forward: plain CG on ``linop(v) = K v + alpha v`` with the same stopping rule as the reference
(``pykeops/pykeops/common/operations.py:90-114``: ``delta = size(b) * eps**2``, stop when
``|r|^2 < delta``); the binder is obtained once and each iteration is one asynchronous map-reduce
launch plus a handful of fused torch vector ops on the same stream.  Backward: a second solve on the
upstream gradient plus one ``Grad_WithSavedForward`` reduction per remaining input
(``torch/operations.py:90-165``).
"""
from __future__ import annotations

import copy

import torch

from ..binder import LoadKeOpsB200
from ..common import axis2cat, complete_aliases, get_optional_flags, get_sizes, get_type
from .genred import GenredAutograd, GenredParameters, _resolve_backend, _set_device


def conjugate_gradient(linop, b, eps=1e-6, max_iter=None, dot=None, numel=None, sync_every=1):
    """Same recurrence and stopping rule as ``ConjugateGradientSolver`` (``common/operations.py:90-114``).

    ``dot(u, v)`` (default: the local sum of products) and ``numel`` (default ``b.numel()``) are
    injectable so that the row-sharded solver (``keops_b200.distributed.ShardedKernelSolve``) can run the
    very same recurrence with all-reduced scalars.

    ``sync_every``: the reference compares ``|r|^2`` with the tolerance on the host after every iteration -- one
    device-to-host synchronisation per map-reduce launch.  With ``sync_every = n > 1`` the test runs ON THE DEVICE: every
    update is masked by an ``active`` flag (an iteration after convergence adds exactly zero), and the host reads the flag
    only every ``n`` iterations.  Iterates and the iteration count are bit-identical to the host-tested loop; at most
    ``n - 1`` launches after convergence are wasted, so this is for small systems whose launch is cheaper than a sync."""
    if dot is None:
        def dot(u, v):
            return (u * v).sum()
    delta = (b.numel() if numel is None else numel) * eps ** 2
    a = torch.zeros_like(b)
    r = b.clone()
    nr2 = dot(r, r)
    conjugate_gradient.last_iterations = 0
    if sync_every > 1:
        return _conjugate_gradient_device_test(linop, a, r, nr2, delta, max_iter, dot, int(sync_every))
    if nr2 < delta:
        return a
    p = r.clone()
    k = 0
    while True:
        Mp = linop(p)
        alp = nr2 / dot(p, Mp)
        a.add_(alp * p)
        r.sub_(alp * Mp)
        nr2new = dot(r, r)
        k += 1
        if nr2new < delta or (max_iter is not None and k >= max_iter):
            break
        p = r + (nr2new / nr2) * p
        nr2 = nr2new
    conjugate_gradient.last_iterations = k
    return a


def _conjugate_gradient_device_test(linop, a, r, nr2, delta, max_iter, dot, sync_every):
    """The loop of :func:`conjugate_gradient` with the convergence test on the device (see there)."""
    active = ~(nr2 < delta)                      # 0-d bool tensor; the reference returns zeros when b is already small
    zero = torch.zeros_like(nr2)
    p = r.clone()
    count = torch.zeros((), dtype=torch.int64, device=r.device)
    k = 0
    while True:
        if k % sync_every == 0 and not bool(active):         # the only host synchronisation
            break
        if max_iter is not None and k >= max_iter:
            break
        Mp = linop(p)
        alp = torch.where(active, nr2 / dot(p, Mp), zero)
        a.add_(alp * p)
        r.sub_(alp * Mp)
        nr2new = dot(r, r)
        count += active
        k += 1
        active = active & ~(nr2new < delta)      # the reference leaves the loop before it updates p
        p = torch.where(active, r + (nr2new / nr2) * p, p)
        nr2 = torch.where(active, nr2new, nr2)
    conjugate_gradient.last_iterations = int(count)
    return a


class KernelSolveAutograd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, *args):
        flags = dict(params.optional_flags)
        flags["multVar_highdim"] = 1 if params.rec_multVar_highdim else 0
        tagCPUGPU, tag1D2D, tagHostDevice = _resolve_backend(params.backend, args)
        nbatchdims = max(len(a.shape) for a in args) - 2
        use_ranges = nbatchdims > 0 or params.ranges is not None
        device_id, device_args = _set_device(params.device_id_request, args)
        binder = LoadKeOpsB200(tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id, params.formula,
                               params.aliases, len(args), params.dtype, "torch", flags).import_module()
        ctx.params, ctx.binder = params, binder
        cargs = tuple(a.detach().contiguous() for a in args)
        varinv = cargs[params.varinvpos]

        def linop(var):
            newargs = cargs[:params.varinvpos] + (var.contiguous(),) + cargs[params.varinvpos + 1:]
            res = binder.genred_pytorch(device_args, params.ranges, params.nx, params.ny, nbatchdims, None, *newargs)
            if params.alpha:
                res += params.alpha * var
            return res

        # small systems: the launch is cheaper than a host synchronisation -- test convergence on the device
        sync_every = 1 if float(params.nx) * float(params.ny) >= 2.5e9 else 8
        result = conjugate_gradient(linop, varinv, params.eps, sync_every=sync_every)
        ctx.save_for_backward(*args, result)
        return result

    @staticmethod
    def backward(ctx, G):
        params, binder = ctx.params, ctx.binder
        args = ctx.saved_tensors[:-1]
        result = ctx.saved_tensors[-1]
        nargs = len(args)
        eta = "Var(%d,%d,%d)" % (nargs, binder.dimout, binder.tagIJ)
        resvar = "Var(%d,%d,%d)" % (nargs + 1, binder.dimout, binder.tagIJ)
        newargs = args[:params.varinvpos] + (G,) + args[params.varinvpos + 1:]
        KinvG = KernelSolveAutograd.apply(params, *newargs)
        grads = []
        for var_ind, sig in enumerate(params.aliases):
            if not ctx.needs_input_grad[var_ind + 1]:
                grads.append(None)
            elif var_ind == params.varinvpos:
                grads.append(KinvG)
            else:
                _, cat, dim, pos = get_type(sig, position_in_list=var_ind)
                var = "Var(%d,%d,%d)" % (pos, dim, cat)
                params_g = copy.copy(params)
                params_g.formula = "Grad_WithSavedForward(%s,%s,%s,%s)" % (params.formula, var, eta, resvar)
                params_g.aliases = list(params.aliases) + [eta, resvar]
                params_g.out = None
                params_g.rec_multVar_highdim = None
                args_g = args[:params.varinvpos] + (result,) + args[params.varinvpos + 1:] + (-KinvG, result)
                grad = GenredAutograd.apply(params_g, *args_g)
                if cat == 2:
                    grad = grad.sum(0).view(-1)
                grads.append(grad.reshape(args[var_ind].shape))
        return (None, *grads)


class KernelSolve:
    """``KernelSolve(formula, aliases, varinvalias, axis=0, ...)(*args, alpha=1e-10, eps=1e-6)`` solves
    ``(alpha Id + K) a = b`` where ``b`` is the argument at ``varinvalias``'s position."""

    def __init__(self, formula, aliases, varinvalias, axis=0, dtype_acc="auto", use_double_acc=False,
                 sum_scheme="auto", enable_chunks=True, rec_multVar_highdim=None, dtype=None, cuda_type=None,
                 use_fast_math=True):
        self.optional_flags = get_optional_flags("Sum", dtype_acc, use_double_acc, sum_scheme, enable_chunks,
                                                 use_fast_math)
        self.formula = "Sum_Reduction(%s,%d)" % (formula, axis2cat(axis))
        self.aliases = complete_aliases(formula, list(aliases))
        if varinvalias[:4] == "Var(":
            self.varinvpos = int(varinvalias[4:varinvalias.find(",")])
        else:
            names = [s[:s.find("=")].strip() for s in self.aliases]
            self.varinvpos = names.index(varinvalias)
        self.rec_multVar_highdim = rec_multVar_highdim
        self.axis = axis

    def __call__(self, *args, backend="auto", device_id=-1, alpha=1e-10, eps=1e-6, ranges=None):
        params = GenredParameters()
        params.formula, params.aliases = self.formula, self.aliases
        params.varinvpos = self.varinvpos
        params.alpha, params.eps = alpha, eps
        params.backend = backend
        params.dtype = str(args[0].dtype).split(".")[1]
        params.device_id_request = device_id
        params.ranges = ranges
        params.optional_flags = self.optional_flags
        params.rec_multVar_highdim = self.rec_multVar_highdim
        params.nx, params.ny = get_sizes(self.aliases, *args)
        params.out = None
        return KernelSolveAutograd.apply(params, *args)
