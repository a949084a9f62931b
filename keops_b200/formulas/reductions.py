"""Reduction descriptors: what is accumulated over the streamed index and what is written out.

Mirrors the semantics (not the code) of ``keopscore/keopscore/formulas/reductions/``:

========================  =======================================  ==============================
reduction                 accumulator (``dimred``)                  output (``dim``)
========================  =======================================  ==============================
Sum                       running sum (dimF)                        sum (dimF)
Max_SumShiftExpWeight     ``(m, s)``, ``m=max F``,                  ``(m, s)`` unnormalised
                          ``s=sum exp(F-m) G`` (1+dimG)             (``Max_SumShiftExpWeight_Reduction.py:16-61``)
Min / Max                 best value (dimF)                         value
ArgMin / ArgMax           (value, index) (2 dimF)                   index, stored as float
Min_ArgMin / Max_ArgMax   (value, index)                            values then indices (2 dimF)
KMin / ArgKMin /          K sorted (value,index) pairs per          K values / K indices / both,
KMin_ArgKMin              component (2 K dimF)                      K-major interleaved
Zero                      --                                        zeros (no kernel is launched)
========================  =======================================  ==============================

``tagI`` follows the reference: ``tagI == 0`` means *output indexed by i, reduction over j*
(``pykeops/pykeops/common/utils.py:11-21``).  Ties in Min/ArgMin/KMin keep the smallest index
(strict ``<``, ``Min_ArgMin_Reduction_Base.py:31-35``, ``KMin_ArgKMin_Reduction.py:71-100``).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace as _dc_replace
from typing import Optional

from . import ir
from .ir import Node, FormulaError
from .autodiff import diff, grad


@dataclass(frozen=True)
class Reduction:
    kind: str                 # "Sum", "LSE", "Min", "Max", "ArgMin", ..., "Zero"
    formula: Node             # F  (for LSE: the scalar formula F)
    tagI: int
    K: int = 1                # for the K-min family
    weight: Optional[Node] = None  # G for LSE
    # provenance, not part of the reduction's identity: ("grad", reduction, variable, gradin, saved forward) when this
    # reduction was produced by Grad_WithSavedForward (lets the binder recognise gradients of known kernels structurally)
    origin: Optional[tuple] = field(default=None, compare=False, repr=False)

    # ---- dimensions ---------------------------------------------------------------------
    @property
    def dimF(self) -> int:
        return self.formula.dim

    @property
    def dim(self) -> int:
        """Output dimension per row."""
        k = self.kind
        if k in ("Sum", "Min", "Max", "ArgMin", "ArgMax", "Zero"):
            return self.dimF
        if k == "LSE":
            return 1 + self.weight.dim
        if k in ("Min_ArgMin", "Max_ArgMax"):
            return 2 * self.dimF
        if k in ("KMin", "ArgKMin"):
            return self.K * self.dimF
        if k == "KMin_ArgKMin":
            return 2 * self.K * self.dimF
        raise FormulaError("unknown reduction " + k)

    @property
    def dimred(self) -> int:
        """Accumulator dimension, as the reference reports it (``get_keops_dll`` tuple)."""
        k = self.kind
        if k in ("Sum", "Min", "Max", "Zero"):
            return self.dimF
        if k == "LSE":
            return 1 + self.weight.dim
        if k in ("ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax"):
            return 2 * self.dimF
        return 2 * self.K * self.dimF

    @property
    def is_zero(self) -> bool:
        return self.kind == "Zero"

    def all_formulas(self):
        return [self.formula] + ([self.weight] if self.weight is not None else [])

    def variables(self):
        vs = set()
        for f in self.all_formulas():
            vs.update(f.variables())
        return sorted(vs, key=lambda v: v.params[0])

    def canonical(self) -> str:
        w = "" if self.weight is None else "," + repr(self.weight)
        return "%s[K=%d,tagI=%d](%r%s)" % (self.kind, self.K, self.tagI, self.formula, w)


def _check_tag(tagI):
    if tagI not in (0, 1):
        raise FormulaError("reduction index tag must be 0 or 1")
    return int(tagI)


def Sum_Reduction(F: Node, tagI: int) -> Reduction:
    if F.is_zero:
        return Zero_Reduction(F.dim, tagI)
    return Reduction("Sum", F, _check_tag(tagI))


def Zero_Reduction(dim: int, tagI: int) -> Reduction:
    return Reduction("Zero", ir.Zero(dim), _check_tag(tagI))


def Max_SumShiftExpWeight_Reduction(F: Node, tagI: int, G: Optional[Node] = None) -> Reduction:
    if F.dim != 1:
        raise FormulaError("Max_SumShiftExpWeight_Reduction requires first formula of dimension 1.")
    if G is None:
        G = ir.IntCst(1)
    return Reduction("LSE", F, _check_tag(tagI), weight=ir.as_node(G))


Max_SumShiftExp_Reduction = Max_SumShiftExpWeight_Reduction


def _simple(kind):
    def make(F: Node, tagI: int) -> Reduction:
        return Reduction(kind, F, _check_tag(tagI))

    make.__name__ = kind + "_Reduction"
    return make


Min_Reduction = _simple("Min")
Max_Reduction = _simple("Max")
ArgMin_Reduction = _simple("ArgMin")
ArgMax_Reduction = _simple("ArgMax")
Min_ArgMin_Reduction = _simple("Min_ArgMin")
Max_ArgMax_Reduction = _simple("Max_ArgMax")


def _kmin(kind):
    def make(F: Node, K: int, tagI: int) -> Reduction:
        K = int(K)
        if K < 1:
            raise FormulaError("K must be a positive integer")
        return Reduction(kind, F, _check_tag(tagI), K=K)

    make.__name__ = kind + "_Reduction"
    return make


KMin_Reduction = _kmin("KMin")
ArgKMin_Reduction = _kmin("ArgKMin")
KMin_ArgKMin_Reduction = _kmin("KMin_ArgKMin")


# ------------------------------------------------------------------------------------------
# gradients of reductions
# ------------------------------------------------------------------------------------------
def Grad_WithSavedForward(red: Reduction, v: Node, gradin: Node, f0: Node) -> Reduction:
    """Gradient reduction of ``red`` w.r.t. ``v``; ``gradin`` is the upstream gradient variable and
    ``f0`` the saved forward output (``autodiff/Grad_WithSavedForward.py``,
    ``reductions/Sum_Reduction.py:42-45``, ``Max_SumShiftExpWeight_Reduction.py:89-103``).

    The new reduction runs over the index *other* than ``v``'s category (``v.cat % 2``): a
    gradient w.r.t. a j-variable is itself a reduction over i.
    """
    if v.op != "Var":
        raise FormulaError("can only differentiate with respect to a Var")
    new_tag = v.params[2] % 2
    k = red.kind
    if k == "Sum":
        return _dc_replace(Sum_Reduction(grad(red.formula, v, gradin), new_tag), origin=("grad", red, v, gradin, f0))
    if k == "LSE":
        F, G = red.formula, red.weight
        M = ir.Extract(f0, 0, 1)
        gradin_S = ir.Extract(gradin, 1, G.dim)
        return Sum_Reduction(grad(ir.Exp(F - M) * G, v, gradin_S), new_tag)
    if k in ("ArgMin", "ArgMax", "ArgKMin", "Zero"):
        return Zero_Reduction(v.dim, new_tag)
    raise NotImplementedError(
        "As of today, KeOps does not support autodiff through the %s reduction." % k
    )


def Diff_WithSavedForward(red: Reduction, v: Node, u: Node, f0: Node) -> Reduction:
    """Forward-mode differential of a reduction (``autodiff/Diff_WithSavedForward.py``; ``Sum_Reduction.py:47-50``,
    ``Max_SumShiftExpWeight_Reduction.py:105-120``): same reduction index as ``red``."""
    k = red.kind
    if k == "Sum":
        return Sum_Reduction(diff(red.formula, v, u), red.tagI)
    if k == "LSE":
        F, G = red.formula, red.weight
        M = ir.Extract(f0, 0, 1)
        return Sum_Reduction(diff(ir.Concat(ir.Zero(1), ir.Exp(F - M) * G), v, u), red.tagI)
    raise NotImplementedError("As of today, KeOps does not support forward-mode autodiff through the %s reduction." % k)


def Diff(red_or_formula, v: Node, u: Node):
    if isinstance(red_or_formula, Reduction):
        return Diff_WithSavedForward(red_or_formula, v, u, None)
    return diff(red_or_formula, v, u)


def Grad(red_or_formula, v: Node, gradin: Node):
    if isinstance(red_or_formula, Reduction):
        return Grad_WithSavedForward(red_or_formula, v, gradin, None)
    return grad(red_or_formula, v, gradin)
