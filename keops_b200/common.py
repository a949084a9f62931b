"""Host-side helpers shared by the torch front-end: alias parsing, option handling and the
pre/post-processing that maps user-level reductions onto engine reductions.

Same behaviour as ``pykeops/pykeops/common/parse_type.py``, ``pykeops/pykeops/common/utils.py:11-36``
and ``pykeops/pykeops/common/operations.py:7-87`` (re-implemented; see the per-function notes).
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

_CATS = {"Vi": 0, "Vj": 1, "Pm": 2}
_SUMLIKE = ("Sum", "Max_SumShiftExp", "Max_SumShiftExpWeight")


def axis2cat(axis: int) -> int:
    """axis is the reduced dimension (python way), cat the remaining one (``common/utils.py:11-21``)."""
    if axis in (0, 1):
        return (axis + 1) % 2
    raise ValueError("Axis should be 0 or 1.")


def cat2axis(cat: int) -> int:
    if cat in (0, 1):
        return (cat + 1) % 2
    raise ValueError("Category should be Vi or Vj.")


def get_type(type_str: str, position_in_list: Optional[int] = None):
    """``"x = Vi(3)"`` / ``"Vj(1,3)"`` / ``"Var(2,3,1)"`` -> (name, cat, dim, pos) (``parse_type.py:71-131``)."""
    s = type_str.replace(" ", "").replace("Vx", "Vi").replace("Vy", "Vj")
    m = re.fullmatch(r"(?:([A-Za-z_][A-Za-z_0-9]*)=)?(Vi|Vj|Pm)\((\d+)(?:,(\d+))?\)", s)
    if m:
        name, kind, a, b = m.groups()
        if b is None:
            return name, _CATS[kind], int(a), position_in_list
        return name, _CATS[kind], int(b), int(a)
    m = re.fullmatch(r"(?:([A-Za-z_][A-Za-z_0-9]*)=)?Var\((\d+),(\d+),(\d+)\)", s)
    if m:
        name, pos, dim, cat = m.groups()
        return name, int(cat), int(dim), int(pos)
    raise ValueError(
        type_str + " type_str does not match the 'var = [Vi|Vj|Pm](dim)' or 'var = [Vi|Vj|Pm](pos,dim) or "
        "'[Vi|Vj|Pm](dim) or '[Vi|Vj|Pm](pos,dim) or Var(pos,dim,cat)'  format: " + type_str)


def complete_aliases(formula: str, aliases: List[str]) -> List[str]:
    """Append the ``Var(k,d,c)`` that appear only inline in the formula (``parse_type.py:9-45``):
    this is how gradient formulas declare the upstream gradient and the saved forward."""
    aliases = list(aliases)
    found = {}
    for m in re.finditer(r"Var\((\d+),(\d+),(\d+)\)", formula.replace(" ", "")):
        found[int(m.group(1))] = m.group(0)
    extra = sorted(p for p in found if p >= len(aliases))
    if extra:
        aliases += [None] * (max(extra) + 1 - len(aliases))
        for p in extra:
            aliases[p] = found[p]
    if any(a is None for a in aliases):
        raise ValueError("[KeOps] variable positions in the formula are not contiguous")
    return aliases


def get_sizes(aliases: Sequence[str], *args) -> Tuple[int, int]:
    nx = ny = None
    for k, sig in enumerate(aliases):
        _, cat, dim, pos = get_type(sig, position_in_list=k)
        if cat == 0 and nx is None:
            nx = args[pos].shape[-2]
        elif cat == 1 and ny is None:
            ny = args[pos].shape[-2]
    return (1 if nx is None else nx), (1 if ny is None else ny)


def get_optional_flags(reduction_op_internal, dtype_acc, use_double_acc, sum_scheme, enable_chunks, use_fast_math):
    """``parse_type.py:134-207``: validation + defaults (block_sum for sum-type reductions)."""
    if dtype_acc != "auto" and use_double_acc:
        raise ValueError("[pyKeOps] you cannot set both options use_double_acc and dtype_acc.")
    if use_double_acc:
        dtype_acc = "float64"
    if dtype_acc != "auto" and reduction_op_internal not in _SUMLIKE:
        raise ValueError("[pyKeOps] parameter dtype_acc should be set to 'auto' for no-sum type reductions (Min, Max, ArgMin, etc.)")
    if sum_scheme == "auto":
        sum_scheme = "block_sum" if reduction_op_internal in _SUMLIKE else "direct_sum"
    if sum_scheme in ("block_sum", "kahan_scheme"):
        if reduction_op_internal not in _SUMLIKE:
            raise ValueError('[pyKeOps] sum_scheme="%s" is only valid for sum type reductions.' % sum_scheme)
    elif sum_scheme != "direct_sum":
        raise ValueError('[pyKeOps] invalid value for option sum_scheme : should be one of "auto", "direct_sum", "block_sum" or "kahan_scheme".')
    return dict(dtype_acc=dtype_acc, sum_scheme=sum_scheme, use_fast_math=1 if use_fast_math else 0,
                enable_chunks=1 if enable_chunks else 0)


def preprocess(reduction_op: str, formula2: Optional[str]):
    """User-level reduction -> engine reduction (``common/operations.py:7-35``)."""
    if reduction_op in ("SumSoftMaxWeight", "SoftMax"):
        return "Max_SumShiftExpWeight", "Concat(IntCst(1)," + formula2 + ")"
    if reduction_op == "LogSumExp":
        return ("Max_SumShiftExpWeight" if formula2 else "Max_SumShiftExp"), formula2
    if reduction_op == "LogSumExpWeight":
        if not formula2:
            raise ValueError("LogSumExpWeight reduction requires a weight formula (given as formula2 kwarg)")
        return "Max_SumShiftExpWeight", formula2
    return reduction_op, formula2


def postprocess(out, reduction_op: str, opt_arg):
    """Engine output -> user output (``common/operations.py:38-87``), torch tensors."""
    if reduction_op in ("SumSoftMaxWeight", "SoftMax"):
        return out[..., 2:] / out[..., 1][..., None]
    if reduction_op in ("ArgMin", "ArgMax"):
        return out.long()
    if reduction_op in ("Min_ArgMin", "MinArgMin", "Max_ArgMax", "MaxArgMax"):
        tmp = out.view(out.shape[:-1] + (2, -1))
        return tmp[..., 0, :], tmp[..., 1, :].long()
    if reduction_op == "KMin":
        out = out.view(out.shape[:-1] + (opt_arg, -1))
        return out.squeeze(-1) if out.shape[-1] == 1 else out
    if reduction_op == "ArgKMin":
        out = out.long().view(out.shape[:-1] + (opt_arg, -1))
        return out.squeeze(-1) if out.shape[-1] == 1 else out
    if reduction_op in ("KMin_ArgKMin", "KMinArgKMin"):
        out = out.view(out.shape[:-1] + (opt_arg, 2, -1))
        vals, idx = out[..., 0, :], out[..., 1, :].long()
        if vals.shape[-1] == 1:
            return vals.squeeze(-1), idx.squeeze(-1)
        return vals, idx
    if reduction_op == "LogSumExp":
        if out.shape[-1] == 2:
            return (out[..., 0] + out[..., 1].log())[..., None]
        return out[..., 0][..., None] + out[..., 1:].log()
    return out
