"""Reverse-mode symbolic differentiation on the vector IR.

``grad(F, v, gradin)`` returns the formula ``[dF/dv]^T . gradin`` with ``dim == v.dim``,
the same contract as the reference's ``Grad(F, v, gradin) = F.DiffT(v, gradin)``
(``keopscore/keopscore/formulas/autodiff/Grad.py:14-27``).  The per-op rules are the usual
calculus identities (cf. each reference op's ``DiffT`` / ``Derivative``, e.g.
``maths/Mult.py:28-38``, ``maths/Divide.py:28-37``, ``maths/Sum.py:26-30``, ``maths/Exp.py:17-19``);
zeros are pruned by the IR constructors so the result stays small.
"""
from __future__ import annotations

from typing import Dict, Tuple

from . import ir
from .ir import Node, FormulaError


def _depends(f: Node, v: Node) -> bool:
    return v in f.variables()


def grad(f: Node, v: Node, gradin: Node) -> Node:
    """Adjoint of the differential of ``f`` w.r.t. the variable ``v`` applied to ``gradin``."""
    if v.op != "Var":
        raise FormulaError("can only differentiate with respect to a Var")
    if gradin.dim != f.dim:
        if gradin.dim == 1:
            gradin = ir.SumT(gradin, f.dim)
        else:
            raise FormulaError(
                "gradient input has dimension %d, formula has dimension %d" % (gradin.dim, f.dim)
            )
    return _grad(f, v, gradin, {})


def diff(f: Node, v: Node, u: Node) -> Node:
    """Forward-mode differential ``D_v f . u`` (dimension ``f.dim``): the adjoint of the gradient, obtained like the
    reference does (``keopscore/keopscore/formulas/autodiff/Diff.py:14-36``): back-propagate a placeholder ``g`` through
    ``f`` -- ``h(g) = [df/dv]^T g`` is linear in ``g`` -- then differentiate ``<h(g), u>`` with respect to ``g``."""
    if v.op != "Var":
        raise FormulaError("can only differentiate with respect to a Var")
    if u.dim != v.dim:
        raise FormulaError("direction has dimension %d, variable has dimension %d" % (u.dim, v.dim))
    used = [w.params[0] for w in f.variables() + u.variables()] + [v.params[0]]
    g = ir.Var(1 + max(used), f.dim, 2)                 # placeholder: it disappears from the result
    h = grad(f, v, g)
    if h.is_zero:
        return ir.Zero(f.dim)
    res = grad(ir.Scalprod(h, u) if v.dim > 1 else ir.Mult(h, u), g, ir.IntCst(1))
    if g in res.variables():
        raise NotImplementedError("Diff: the placeholder did not cancel (an operation's gradient rule is not linear in gradin)")
    return res


def divergence(f: Node, v: Node) -> Node:
    """``sum_k d f_k / d v_k`` -- the trace of the Jacobian (``keopscore/keopscore/formulas/autodiff/Divergence.py``):
    sum over k of component k of the differential applied to the k-th basis vector."""
    if v.op != "Var":
        raise FormulaError("can only differentiate with respect to a Var")
    if f.dim != v.dim:
        raise FormulaError("Divergence requires formula and variable to have the same dimension.")
    total = None
    for k in range(v.dim):
        e_k = ir.ElemT(ir.IntCst(1), v.dim, k)
        term = ir.Elem(diff(f, v, e_k), k)
        total = term if total is None else ir.Add(total, term)
    return total


def grad_matrix(f: Node, v: Node) -> Node:
    """Jacobian of ``f`` w.r.t. ``v`` as a vector of dimension ``f.dim * v.dim``: the gradients of the components of ``f``,
    concatenated (``keopscore/keopscore/formulas/maths/GradMatrix.py:14-28``)."""
    res = None
    for i in range(f.dim):
        g_i = grad(f, v, ir.ElemT(ir.IntCst(1), f.dim, i))
        if g_i.is_zero:
            g_i = ir.SumT(ir.IntCst(0), v.dim) if v.dim > 1 else ir.IntCst(0)
        res = g_i if res is None else ir.Concat(res, g_i)
    return res


def laplacian(f: Node, v: Node) -> Node:
    """``Divergence(Grad(f, v, 1), v)`` for a scalar formula (``autodiff/Laplacian.py``)."""
    if f.dim != 1:
        raise FormulaError("Laplacian is only implemented for scalar formula.")
    return divergence(grad(f, v, ir.IntCst(1)), v)


_LINEAR_IF_ALL = frozenset(("Add", "Subtract", "Minus", "Sum", "SumT", "Extract", "ExtractT", "Concat"))
_LINEAR_IF_ONE = frozenset(("Mult", "Scalprod", "MatVecMult", "VecMatMult", "TensorProd", "TensorDot"))


def is_linear(f: Node, v: Node) -> bool:
    """Structural linearity of ``f`` in the variable ``v``, by the reference's rules (``formulas/Operation.py:16-27`` and the
    ``linearity_type`` of each operation: sums need every term linear, products exactly one factor, quotients the numerator)."""
    if f.op == "Var":
        return f is v
    if f.op == "Zero":
        return True
    if f.op in _LINEAR_IF_ALL:
        return all(is_linear(a, v) for a in f.args)
    if f.op in _LINEAR_IF_ONE:
        return sum(is_linear(a, v) for a in f.args) == 1
    if f.op == "Divide":
        return is_linear(f.args[0], v)
    return False


def adjoint_operator(f: Node, v: Node, u: Node) -> Node:
    """``AdjointOperator(f, v, u)``: for ``f`` linear in ``v`` (matrix ``M``, ``f.dim x v.dim``), the vector ``M^T u``
    (``keopscore/keopscore/formulas/LinearOperators/AdjointOperator.py``).  The adjoint of a linear map is its transposed
    Jacobian applied to ``u``: the reverse-mode gradient."""
    if v.op != "Var":
        raise FormulaError("can only take the adjoint with respect to a Var")
    if not is_linear(f, v):
        raise FormulaError("Formula is not linear with respect to variable.")
    if u.dim != f.dim:
        raise FormulaError("AdjointOperator: the argument has dimension %d, the formula has dimension %d" % (u.dim, f.dim))
    return grad(f, v, u)


def trace_operator(f: Node, v: Node, w: Node = None) -> Node:
    """``TraceOperator(f, v, w=1)`` for ``f`` linear in ``v`` with matrix ``M`` (``LinearOperators/TraceOperator.py:30-52``):
    ``f.dim == v.dim``: ``tr(diag(w) M)`` (``w * tr(M)`` for a scalar ``w``); ``f.dim == 1`` (``M`` is a row): ``sum_i w_i M_i``.
    Entries of ``M`` are read off the transposed Jacobian: ``M_kk = [M^T e_k]_k``, ``M_i = [M^T 1]_i``."""
    if v.op != "Var":
        raise FormulaError("can only take the trace with respect to a Var")
    if not is_linear(f, v):
        raise FormulaError("Formula is not linear with respect to variable.")
    d = v.dim
    wdim = 1 if w is None else w.dim
    if wdim not in (1, d):
        raise FormulaError("Dimension of argument w should be 1 or equal dimension of v.")
    if f.dim not in (1, d):
        raise FormulaError("Output dimension should be 1 or %d (%s, dim %d)." % (d, f.op, f.dim))
    if f.dim == 1:
        row = grad(f, v, ir.IntCst(1))                                   # M^T 1 = the row itself
        if w is None:
            return ir.Sum(row)
        return ir.Mult(w, ir.Sum(row)) if wdim == 1 and d > 1 else ir.Sum(ir.Mult(w, row))
    total = None
    for k in range(d):
        m_kk = ir.Elem(grad(f, v, ir.ElemT(ir.IntCst(1), d, k)), k)
        if w is not None and wdim == d:
            m_kk = ir.Mult(ir.Elem(w, k), m_kk)
        total = m_kk if total is None else ir.Add(total, m_kk)
    return ir.Mult(w, total) if (w is not None and wdim == 1) else total



def _grad(f: Node, v: Node, g: Node, memo: Dict[Tuple[int, int], Node]) -> Node:
    if g.is_zero or not _depends(f, v):
        return ir.Zero(v.dim)
    key = (id(f), id(g))
    hit = memo.get(key)
    if hit is not None:
        return hit
    res = _rule(f, v, g, memo)
    memo[key] = res
    return res


def _reduce_to(g: Node, dim: int) -> Node:
    """Adjoint of broadcasting a ``dim``-vector up to ``g.dim``."""
    if g.dim == dim:
        return g
    if dim == 1:
        return ir.Sum(g)
    raise FormulaError("internal: cannot reduce gradient of dim %d to %d" % (g.dim, dim))


def _rule(f: Node, v: Node, g: Node, memo) -> Node:
    op = f.op
    G = lambda child, gi: _grad(child, v, _reduce_to(gi, child.dim), memo)

    if op == "Var":
        return g if f is v else ir.Zero(v.dim)
    if op in ("Zero", "IntCst", "RatCst", "FloatCst"):
        return ir.Zero(v.dim)

    if op == "Add":
        a, b = f.args
        return G(a, g) + G(b, g)
    if op == "Subtract":
        a, b = f.args
        return G(a, g) - G(b, g)
    if op == "Minus":
        (a,) = f.args
        return G(a, -g)
    if op == "Mult":
        a, b = f.args
        return G(a, g * b) + G(b, g * a)
    if op == "Divide":
        a, b = f.args
        return G(a, g / b) - G(b, (g * a) / ir.Square(b))
    if op == "Square":
        (a,) = f.args
        return G(a, ir.IntCst(2) * (a * g))
    if op == "Inv":
        (a,) = f.args
        return G(a, -(g / ir.Square(a)))
    if op == "Exp":
        (a,) = f.args
        return G(a, g * f)
    if op == "Log":
        (a,) = f.args
        return G(a, g / a)
    if op == "Sqrt":
        (a,) = f.args
        return G(a, (ir.RatCst(1, 2) * ir.Rsqrt(a)) * g)
    if op == "Rsqrt":
        (a,) = f.args
        return G(a, (ir.RatCst(-1, 2) * ir.Pow(f, 3)) * g)
    if op == "Pow":
        (a,) = f.args
        (m,) = f.params
        return G(a, (ir.IntCst(m) * ir.Pow(a, m - 1)) * g)
    if op == "Powf":
        a, b = f.args
        return G(a, (b * ir.Powf(a, b - ir.IntCst(1))) * g) + G(b, (ir.Log(a) * f) * g)
    if op == "Sin":
        (a,) = f.args
        return G(a, ir.Cos(a) * g)
    if op == "Cos":
        (a,) = f.args
        return G(a, -(ir.Sin(a) * g))
    if op == "Acos":
        (a,) = f.args
        return G(a, -(ir.Rsqrt(ir.IntCst(1) - ir.Square(a)) * g))
    if op == "Asin":
        (a,) = f.args
        return G(a, ir.Rsqrt(ir.IntCst(1) - ir.Square(a)) * g)
    if op == "Atan":
        (a,) = f.args
        return G(a, g / (ir.IntCst(1) + ir.Square(a)))
    if op == "Atan2":
        a, b = f.args  # atan2(a, b)
        den = ir.Square(a) + ir.Square(b)
        return G(a, (b / den) * g) + G(b, -((a / den) * g))
    if op == "Abs":
        (a,) = f.args
        return G(a, ir.Sign(a) * g)
    if op == "ReLU":
        (a,) = f.args
        return G(a, ir.Step(a) * g)
    if op == "XLogX":
        (a,) = f.args
        return G(a, (ir.Log(a) + ir.IntCst(1)) * g)
    if op == "SinXDivX":
        (a,) = f.args
        return G(a, (ir.Cos(a) / a - ir.Sin(a) / ir.Square(a)) * g)
    if op in ("Sign", "Step", "Floor", "Round", "LessThan", "LessOrEqual", "Equal", "NotEqual",
              "ArgMax", "ArgMin", "OneHot", "DiffClampInt"):
        return ir.Zero(v.dim)
    if op == "IfElse":
        c, a, b = f.args
        z = ir.Zero(g.dim)
        return G(a, ir.IfElse(c, g, z)) + G(b, ir.IfElse(c, z, g))
    if op == "ClampInt":
        (a,) = f.args
        lo, hi = f.params
        return G(a, ir.DiffClampInt(a, lo, hi) * g)
    if op == "Clamp":
        x, lo, hi = f.args
        # same alternative form as the reference (``maths/Clamp.py``)
        alt = ir.ClampInt((x - lo) / (hi - lo), 0, 1) * (hi - lo) + lo
        return _grad(alt, v, g, memo)
    if op == "Mod":
        x, n, d = f.args
        # the reference differentiates the alternative form x - n*Floor((x-d)/n) (``maths/Mod.py:14-20``), which also
        # carries the derivative with respect to the modulus n: -Floor((x-d)/n) * g
        alt = x - n * ir.Floor((x - d) / n)
        return _grad(alt, v, g, memo)

    if op == "Sum":
        (a,) = f.args
        return G(a, ir.SumT(g, a.dim))
    if op == "SumT":
        (a,) = f.args
        return G(a, ir.Sum(g))
    if op == "Scalprod":
        a, b = f.args
        return G(a, g * b) + G(b, g * a)
    if op == "Max" or op == "Min":
        (a,) = f.args
        arg = ir.ArgMax(a) if op == "Max" else ir.ArgMin(a)
        return G(a, ir.OneHot(arg, a.dim) * g)
    if op == "Concat":
        a, b = f.args
        return G(a, ir.Extract(g, 0, a.dim)) + G(b, ir.Extract(g, a.dim, b.dim))
    if op == "Extract":
        (a,) = f.args
        start, _ = f.params
        return G(a, ir.ExtractT(g, start, a.dim))
    if op == "ExtractT":
        (a,) = f.args
        start, _ = f.params
        return G(a, ir.Extract(g, start, a.dim))
    if op == "MatVecMult":
        A, b = f.args
        return G(A, ir.TensorProd(g, b)) + G(b, ir.VecMatMult(g, A))
    if op == "VecMatMult":
        b, A = f.args
        return G(A, ir.TensorProd(b, g)) + G(b, ir.MatVecMult(A, g))
    if op == "TensorProd":
        a, b = f.args
        return G(a, ir.MatVecMult(g, b)) + G(b, ir.VecMatMult(a, g))
    # ---- complex arithmetic (keopscore/keopscore/formulas/complex/*.py).  The reference adds the two partial gradients of a
    # binary complex operation with ComplexAdd, which refuses odd sizes (a real variable of size 1 or 3); they are gradients
    # with respect to the same variable, of the same size, so a plain sum is the same thing wherever the reference works.
    if op == "ComplexReal":
        return G(f.args[0], ir.Real2Complex(g))
    if op == "ComplexImag":
        return G(f.args[0], ir.Imag2Complex(g))
    if op == "Real2Complex":
        return G(f.args[0], ir.ComplexReal(g))
    if op == "Imag2Complex":
        return G(f.args[0], ir.ComplexImag(g))
    if op == "Conj":
        return G(f.args[0], ir.Conj(g))
    if op == "ComplexSquareAbs":
        (a,) = f.args
        return _grad(ir.ComplexReal(ir.ComplexMult(a, ir.Conj(a))), v, g, memo)
    if op == "ComplexMult":
        a, b = f.args
        ga, gb = ir.ComplexMult(ir.Conj(b), g), ir.ComplexMult(ir.Conj(a), g)
        if a.dim == 2 and b.dim > 2:
            ga = ir.ComplexSum(ga)
        elif b.dim == 2 and a.dim > 2:
            gb = ir.ComplexSum(gb)
        return G(a, ga) + G(b, gb)
    if op in ("ComplexAdd", "ComplexSubtract"):
        a, b = f.args
        ga, gb = g, g
        if a.dim == 2 and b.dim > 2:
            ga = ir.ComplexSum(g)
        elif b.dim == 2 and a.dim > 2:
            gb = ir.ComplexSum(g)
        return G(a, ga) + G(b, gb) if op == "ComplexAdd" else G(a, ga) - G(b, gb)
    if op == "ComplexRealScal":
        a, b = f.args
        return _grad(ir.ComplexMult(ir.Real2Complex(a), b), v, g, memo)
    if op == "ComplexExp":
        (a,) = f.args
        mod = ir.Exp(ir.ComplexReal(a))
        alt = ir.Real2Complex(mod * ir.Cos(ir.ComplexImag(a))) + ir.Imag2Complex(mod * ir.Sin(ir.ComplexImag(a)))
        return _grad(alt, v, g, memo)
    if op == "ComplexExp1j":
        (a,) = f.args
        return _grad(ir.Real2Complex(ir.Cos(a)) + ir.Imag2Complex(ir.Sin(a)), v, g, memo)
    if op == "ComplexSum":
        (a,) = f.args
        return G(a, ir.ComplexSumT(g, a.dim))
    if op == "ComplexSumT":
        return G(f.args[0], ir.ComplexSum(g))
    if op == "BSpline":
        raise NotImplementedError("KeOps BSplines are not yet differentiable. Please consider disabling autograd for the "
                                  "relevant variables with 'x.detach()'.")       # maths/BSpline.py:97-103
    if op == "SoftDTW_SqDist":
        x, y, gamma = f.args
        if v in gamma.variables():
            raise FormulaError("autograd wrt gamma in SoftDTW_SqDist operation not implemented.")
        full = ir.GradSoftDTW_SqDist(x, y, gamma) * g                            # SoftDTW_SqDist.py:95-106
        return G(x, ir.Extract(full, 0, x.dim)) + G(y, ir.Extract(full, x.dim, y.dim))
    if op == "GradSoftDTW_SqDist":
        raise FormulaError("autograd for GradSoftDTW_SqDist operation not implemented.")
    if op == "TensorDot":
        # the adjoint of a contraction is the contraction of the incoming gradient with the other operand over that
        # operand's kept axes, stored back in the layout of the differentiated operand (``maths/TensorDot.py:189-216``)
        a, b = f.args
        lay = ir.TensorDotLayout(*f.params)
        return G(a, ir.TensorDot(g, b, *lay.adjoint(0))) + G(b, ir.TensorDot(g, a, *lay.adjoint(1)))
    raise NotImplementedError("no differentiation rule for operation " + op)
