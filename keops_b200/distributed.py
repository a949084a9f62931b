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


class ShardedGenred:
    """``ShardedGenred(formula, aliases, reduction_op, axis, ...)(*local_args, n_streamed=N, mode=...)``.

    ``local_args``: output-side variables = this rank's row slab; streamed variables = this rank's slab
    (``mode="all_gather"``), the full array on ``src`` and anything/None elsewhere (``mode="broadcast"``),
    or the full array everywhere (``mode="replicated"``); parameters = full.  Returns this rank's slab
    of the output.  ``compute`` (default: the CUDA ``Genred``) is injectable so that the sharding and
    collective logic can be tested on CPU with the ``gloo`` backend."""

    def __init__(self, formula: str, aliases: Sequence[str], reduction_op: str = "Sum", axis: int = 0,
                 group=None, compute: Optional[Callable] = None, **genred_kwargs):
        self.formula, self.reduction_op, self.axis = formula, reduction_op, axis
        self.group = group
        red_internal, formula2 = preprocess(reduction_op, genred_kwargs.get("formula2"))
        probe = "%s_Reduction(%s,%d%s)" % (red_internal, formula, axis2cat(axis), ",%s" % formula2 if formula2 else "")
        self.aliases = complete_aliases(probe, list(aliases))
        self.cats, self.dims = [], []
        for k, sig in enumerate(self.aliases):
            _, cat, dim, pos = get_type(sig, position_in_list=k)
            if pos != k:
                raise ValueError("ShardedGenred expects aliases in positional order")
            self.cats.append(cat)
            self.dims.append(dim)
        self.out_cat = axis2cat(axis)          # category of the output index
        if compute is None:
            from .torch import Genred
            compute = Genred(formula, list(aliases), reduction_op=reduction_op, axis=axis, **genred_kwargs)
        self.compute = compute

    def __call__(self, *local_args, n_streamed: Optional[int] = None, mode: str = "all_gather", src: int = 0,
                 **call_kwargs):
        if mode not in ("all_gather", "broadcast", "replicated"):
            raise ValueError("mode must be 'all_gather', 'broadcast' or 'replicated'")
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        args: List[torch.Tensor] = []
        ref = next(a for a, c in zip(local_args, self.cats) if c == self.out_cat and a is not None)
        for k, (a, cat) in enumerate(zip(local_args, self.cats)):
            if cat == self.out_cat or cat == 2 or world == 1 or mode == "replicated":
                args.append(a)
                continue
            if n_streamed is None:
                raise ValueError("n_streamed (global size of the reduced index) is required")
            if mode == "all_gather":
                args.append(all_gather_rows(a, n_streamed, self.group))
            else:
                if rank == src:
                    buf = a.contiguous()
                else:
                    buf = torch.empty((n_streamed, self.dims[k]), dtype=ref.dtype, device=ref.device)
                dist.broadcast(buf, src=dist.get_global_rank(self.group, src) if self.group is not None else src,
                               group=self.group)
                args.append(buf)
        return self.compute(*args, **call_kwargs)


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
