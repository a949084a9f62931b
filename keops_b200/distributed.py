"""Multi-GPU execution of a reduction by sharding the *output index* (one process per GPU).

The reference has no distributed layer; its only multi-GPU code is a thread-pool example
(``pykeops/pykeops/examples/pytorch/plot_multi_gpu.py:83-96``, "shard i, concatenate").  Output rows
are independent (``keopscore/keopscore/mapreduce/gpu/GpuReduc1D.py:50,91-93``), so the partitioning is:

* output-side variables (category = the output's index) and the output itself are row-sharded:
  rank r owns rows ``shard_bounds(n, world, r)``;
* streamed variables (the reduced index) must be complete on every rank: they are replicated with
  one collective per variable per call -- ``all_gather`` when every rank holds its own slab,
  ``broadcast`` when one rank holds the full array -- over NCCL (NVLink / NVSwitch);
* parameters are replicated scalars; there is **no** cross-GPU reduction and arg-type outputs need no
  index fix-up (the replica is complete, so indices are global).

For a gradient with respect to a streamed variable, shard by *that* variable's index instead
(``Grad_WithSavedForward`` w.r.t. a j-variable is itself a reduction over i with output indexed by j,
``keopscore/keopscore/formulas/reductions/Sum_Reduction.py:42-45``): call ``ShardedGenred`` on the
gradient formula with ``axis`` flipped.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .common import complete_aliases, get_type, preprocess, axis2cat


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced row slab of rank ``rank``: sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate row slabs of (possibly) unequal heights from all ranks, in rank order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    if local.shape[0] != sizes[rank][1] - sizes[rank][0]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local.shape[0], sizes[rank][1] - sizes[rank][0]))
    hmax = max(b - a for a, b in sizes)
    full = torch.empty((world * hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.shape[0] == hmax:
        padded = local.contiguous()
    else:
        padded = torch.zeros((hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    dist.all_gather_into_tensor(full, padded, group=group)
    if all(b - a == hmax for a, b in sizes):
        return full
    return torch.cat([full[r * hmax: r * hmax + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


def _replicate(a, k, cat_dims, n_total, mode, src, ref, group):
    """One streamed operand made complete on every rank: all_gather of row slabs / broadcast from ``src`` / already there."""
    if mode == "replicated":
        return a
    if n_total is None:
        raise ValueError("the global size of the gathered index is required")
    if mode == "all_gather":
        return all_gather_rows(a, n_total, group)
    if mode != "broadcast":
        raise ValueError("mode must be 'all_gather', 'broadcast' or 'replicated'")
    rank = dist.get_rank(group)
    if rank == src:
        buf = a.contiguous()
    else:
        buf = torch.empty((n_total, cat_dims[k]), dtype=ref.dtype, device=ref.device)
    dist.broadcast(buf, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return buf


class ShardedGenred:
    """``ShardedGenred(formula, aliases, reduction_op, axis, ...)(*local_args, n_streamed=N, mode=...)``.

    ``local_args``: output-side variables = this rank's row slab; streamed variables = this rank's slab
    (``mode="all_gather"``), the full array on ``src`` and anything/None elsewhere (``mode="broadcast"``),
    or the full array everywhere (``mode="replicated"``); parameters = full.  ``mode`` may also be a dict
    ``{argument position: mode}`` (other streamed arguments: ``"all_gather"``), e.g. a point cloud that was replicated
    once and a vector that changes every call.  Returns this rank's slab
    of the output.  ``compute`` (default: the CUDA ``Genred``) is injectable so that the sharding and
    collective logic can be tested on CPU with the ``gloo`` backend.

    ``gradient(var)`` returns the reduction the reference's backward would launch for that input
    (``Grad_WithSavedForward``, ``pykeops/pykeops/torch/generic/generic_red.py:190-214``), sharded by the *differentiated
    variable's own index*: a gradient w.r.t. a streamed variable ``y_j`` is a reduction over ``i`` whose output is indexed by
    ``j`` (``keopscore/keopscore/formulas/reductions/Sum_Reduction.py:42-45``), so every rank produces its own slab of
    ``d/dy`` from its slab of ``y`` and the gathered ``i``-side arrays -- still no cross-GPU reduction."""

    def __init__(self, formula: str, aliases: Sequence[str], reduction_op: str = "Sum", axis: int = 0,
                 group=None, compute: Optional[Callable] = None, **genred_kwargs):
        self.formula, self.reduction_op, self.axis = formula, reduction_op, axis
        self.group = group
        red_internal, formula2 = preprocess(reduction_op, genred_kwargs.get("formula2"))
        opt_arg = genred_kwargs.get("opt_arg")
        self.red_formula = "%s_Reduction(%s%s,%d%s)" % (red_internal, formula, ",%s" % opt_arg if opt_arg else "",
                                                       axis2cat(axis), ",%s" % formula2 if formula2 else "")
        self.aliases = complete_aliases(self.red_formula, list(aliases))
        self.cats, self.dims = [], []
        for k, sig in enumerate(self.aliases):
            _, cat, dim, pos = get_type(sig, position_in_list=k)
            if pos != k:
                raise ValueError("ShardedGenred expects aliases in positional order")
            self.cats.append(cat)
            self.dims.append(dim)
        self.out_cat = axis2cat(axis)          # category of the output index
        self._genred_kwargs = dict(genred_kwargs)
        self._user_compute = compute is not None
        if compute is None:
            from .torch import Genred
            compute = Genred(formula, list(aliases), reduction_op=reduction_op, axis=axis, **genred_kwargs)
        self.compute = compute

    def _mode_of(self, mode, k):
        if isinstance(mode, dict):
            return mode.get(k, "all_gather")
        return mode

    def __call__(self, *local_args, n_streamed: Optional[int] = None, mode="all_gather", src: int = 0,
                 **call_kwargs):
        if not isinstance(mode, dict) and mode not in ("all_gather", "broadcast", "replicated"):
            raise ValueError("mode must be 'all_gather', 'broadcast' or 'replicated'")
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        args: List[torch.Tensor] = []
        ref = next(a for a, c in zip(local_args, self.cats) if c == self.out_cat and a is not None)
        for k, (a, cat) in enumerate(zip(local_args, self.cats)):
            if cat == self.out_cat or cat == 2 or world == 1:
                args.append(a)
            else:
                args.append(_replicate(a, k, self.dims, n_streamed, self._mode_of(mode, k), src, ref, self.group))
        return self.compute(*args, **call_kwargs)

    def gradient(self, var, compute: Optional[Callable] = None) -> "ShardedGradient":
        """``var``: position (or alias name) of the input to differentiate with respect to."""
        if isinstance(var, str):
            var = [s[:s.find("=")].strip() for s in self.aliases].index(var)
        return ShardedGradient(self, int(var), compute)


class ShardedGradient:
    """``g = ShardedGenred(...).gradient(var)``; ``g(*local_args, eta_local, fwd_local, n_i=M, n_j=N)`` -> this rank's slab of
    the gradient of ``<eta, Reduction(...)>`` with respect to input ``var``.

    Every array is row-sharded by its OWN index (``i``-indexed arrays over M, ``j``-indexed ones over N; ``eta`` and the
    saved forward ``fwd`` follow the forward's output index).  The gradient reduction's output is indexed like ``var``:
    arrays of that category stay local, the others are all-gathered (``mode``: as for :class:`ShardedGenred`, the positions
    ``nargs`` / ``nargs + 1`` standing for ``eta`` / ``fwd``).  A gradient w.r.t. a parameter is the sum over all output rows:
    local reduction + one all-reduce of ``dim`` numbers (the reference sums in Python, ``generic_red.py:216-235``)."""

    def __init__(self, parent: ShardedGenred, var: int, compute: Optional[Callable] = None):
        from .formulas.parser import parse_reduction
        from .binder import _normalise_aliases
        self.parent, self.var = parent, var
        red = parse_reduction(parent.red_formula, _normalise_aliases(parent.aliases))
        self.dimout, self.tagIJ = red.dim, red.tagI
        nargs = len(parent.aliases)
        cat, dim = parent.cats[var], parent.dims[var]
        self.var_cat = cat
        eta = "Var(%d,%d,%d)" % (nargs, self.dimout, self.tagIJ)
        res = "Var(%d,%d,%d)" % (nargs + 1, self.dimout, self.tagIJ)
        self.red_formula = "Grad_WithSavedForward(%s,Var(%d,%d,%d),%s,%s)" % (parent.red_formula, var, dim, cat, eta, res)
        self.aliases = list(parent.aliases) + [eta, res]
        self.cats = list(parent.cats) + [self.tagIJ, self.tagIJ]
        self.dims = list(parent.dims) + [self.dimout, self.dimout]
        # a parameter's gradient is reduced like the forward (output indexed by the forward's index), then summed over rows
        self.out_cat = cat if cat in (0, 1) else self.tagIJ
        if compute is None:
            from .torch.genred import RawReduction
            from .common import get_optional_flags
            kw = parent._genred_kwargs
            red_internal, _ = preprocess(parent.reduction_op, kw.get("formula2"))
            flags = get_optional_flags(red_internal, kw.get("dtype_acc", "auto"), kw.get("use_double_acc", False),
                                       kw.get("sum_scheme", "auto"), kw.get("enable_chunks", True), kw.get("use_fast_math", True))
            compute = RawReduction(self.red_formula, self.aliases, flags)
        self.compute = compute

    def __call__(self, *local_args, n_i: Optional[int] = None, n_j: Optional[int] = None, mode="all_gather", src: int = 0,
                 **call_kwargs):
        if len(local_args) != len(self.aliases):
            raise ValueError("expected the forward's %d arguments followed by eta and the saved forward" % (len(self.aliases) - 2))
        group = self.parent.group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        n_of = {0: n_i, 1: n_j}
        ref = next(a for a, c in zip(local_args, self.cats) if c == self.out_cat and a is not None)
        args: List[torch.Tensor] = []
        for k, (a, cat) in enumerate(zip(local_args, self.cats)):
            if cat == self.out_cat or cat == 2 or world == 1:
                args.append(a)
            else:
                args.append(_replicate(a, k, self.dims, n_of[cat], self.parent._mode_of(mode, k), src, ref, group))
        g = self.compute(*args, **call_kwargs)
        if self.var_cat == 2:
            g = (1.0 * g).sum(-2)
            if world > 1:
                dist.all_reduce(g, group=group)
            g = g.reshape(local_args[self.var].shape)
        return g


class ShardedKernelSolve:
    """Row-sharded conjugate gradient for ``(alpha Id + K) a = b`` (single-process solver:
    ``keops_b200.torch.KernelSolve``, reference ``pykeops/pykeops/torch/operations.py:29-87`` and
    ``pykeops/pykeops/common/operations.py:90-114``).

    ``K_ij = formula(x_i, y_j, ...)`` is square (M = N = ``n_total``).  Every rank owns one contiguous row
    slab of the CG vectors (``b``, ``a``, ``r``, ``p``) and of the output-side variables.  The fixed streamed
    variables are replicated ONCE per solve; each iteration then costs one ``all_gather`` of the search
    direction ``p`` (4 MB at N = 10^6), one local map-reduce launch (this rank's rows against the full
    ``p``) and two all-reduced scalars (``p.Kp`` and ``|r|^2``).  The recurrence and the stopping rule
    are the reference's, so every rank takes the same number of iterations.

    ``local_args`` follow the aliases: output-side variables and the right-hand side (at ``varinvalias``'s
    position) = this rank's slab; other streamed variables = this rank's slab (``mode="all_gather"``) or
    the full array (``mode="replicated"``); parameters = full.  Returns this rank's slab of ``a``.
    ``compute(*full_args) -> local rows of K v`` is injectable (CPU ``gloo`` tests)."""

    def __init__(self, formula: str, aliases: Sequence[str], varinvalias: str, axis: int = 0, group=None,
                 compute: Optional[Callable] = None, **genred_kwargs):
        self.group = group
        probe = "Sum_Reduction(%s,%d)" % (formula, axis2cat(axis))
        self.aliases = complete_aliases(probe, list(aliases))
        self.cats, self.dims = [], []
        for k, sig in enumerate(self.aliases):
            _, cat, dim, pos = get_type(sig, position_in_list=k)
            if pos != k:
                raise ValueError("ShardedKernelSolve expects aliases in positional order")
            self.cats.append(cat)
            self.dims.append(dim)
        if varinvalias[:4] == "Var(":
            self.varinvpos = int(varinvalias[4:varinvalias.find(",")])
        else:
            self.varinvpos = [s[:s.find("=")].strip() for s in self.aliases].index(varinvalias)
        self.out_cat = axis2cat(axis)
        if self.cats[self.varinvpos] != 1 - self.out_cat:
            raise ValueError("the unknown of KernelSolve must be indexed by the reduced index")
        if compute is None:
            from .torch import Genred
            compute = Genred(formula, list(aliases), reduction_op="Sum", axis=axis, **genred_kwargs)
        self.compute = compute
        self.last_iterations = 0

    def __call__(self, *local_args, n_total: int, alpha: float = 1e-10, eps: float = 1e-6, mode: str = "all_gather",
                 max_iter: Optional[int] = None):
        from .torch.kernel_solve import conjugate_gradient
        if mode not in ("all_gather", "replicated"):
            raise ValueError("mode must be 'all_gather' or 'replicated'")
        on = dist.is_initialized()
        world = dist.get_world_size(self.group) if on else 1
        fixed: List[Optional[torch.Tensor]] = []
        for k, (a, cat) in enumerate(zip(local_args, self.cats)):
            if k == self.varinvpos:
                fixed.append(None)
            elif cat == self.out_cat or cat == 2 or world == 1 or mode == "replicated":
                fixed.append(a)
            else:
                fixed.append(all_gather_rows(a, n_total, self.group))     # once per solve
        b = local_args[self.varinvpos]

        def linop(p_local):
            p_full = all_gather_rows(p_local.contiguous(), n_total, self.group) if world > 1 else p_local
            args = [p_full if k == self.varinvpos else a for k, a in enumerate(fixed)]
            res = self.compute(*args)
            if alpha:
                res = res + alpha * p_local
            return res

        def dot(u, v):
            s = (u * v).sum()
            if world > 1:
                dist.all_reduce(s, group=self.group)
            return s

        a = conjugate_gradient(linop, b, eps=eps, max_iter=max_iter, dot=dot, numel=int(n_total) * b[0].numel())
        self.last_iterations = conjugate_gradient.last_iterations
        return a
