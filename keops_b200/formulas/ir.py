"""Vector-level formula IR.

A formula ``F(x_i, y_j, p)`` is a DAG of immutable, hash-consed :class:`Node` objects.  Every
node has a static output dimension ``dim``; element-wise operations broadcast a ``dim == 1``
operand against a ``dim == n`` operand exactly like the reference's ``VectorizedScalarOp``
(``keopscore/keopscore/formulas/VectorizedScalarOp.py:7-24``).

Unlike the reference -- where each ``Operation`` prints its own C++ block
(``keopscore/keopscore/formulas/Operation.py:66-116``) -- nodes here carry no code: the vector DAG
is lowered to a *scalar* DAG by :mod:`keops_b200.codegen.lower`, which is where simplification,
CSE, FMA fusion and i/j/parameter hoisting happen.  Only structural simplifications that keep
autodiff output small (zero / one propagation) are applied at construction time.
"""
from __future__ import annotations

from fractions import Fraction
from typing import Dict, List, Optional, Tuple


class FormulaError(ValueError):
    """Raised for malformed formulas (mirrors ``KeOps_Error`` -> ``ValueError``,
    ``keopscore/keopscore/utils/misc_utils.py:28-35``)."""


_INTERN: Dict[tuple, "Node"] = {}


class Node:
    __slots__ = ("op", "args", "params", "dim", "_key", "_vars")

    def __new__(cls, op: str, args: Tuple["Node", ...] = (), params: tuple = (), dim: int = 1):
        key = (op, tuple(id(a) for a in args), params, dim)
        hit = _INTERN.get(key)
        if hit is not None:
            return hit
        self = object.__new__(cls)
        self.op = op
        self.args = tuple(args)
        self.params = tuple(params)
        self.dim = int(dim)
        self._key = key
        self._vars = None
        _INTERN[key] = self
        return self

    # ---- structural helpers -------------------------------------------------------------
    def variables(self) -> List["Node"]:
        """All ``Var`` leaves, sorted by index (cf. ``Operation.Vars_``, ``Operation.py:36-39``)."""
        if self._vars is None:
            if self.op == "Var":
                vs = {self}
            else:
                vs = set()
                for a in self.args:
                    vs.update(a.variables())
            self._vars = sorted(vs, key=lambda v: v.params[0])
        return self._vars

    @property
    def is_zero(self) -> bool:
        return self.op == "Zero"

    def const_value(self) -> Optional[Fraction]:
        if self.op == "IntCst":
            return Fraction(self.params[0])
        if self.op == "RatCst":
            return Fraction(self.params[0], self.params[1])
        return None

    def __repr__(self):
        if self.op == "Var":
            return "Var(%d,%d,%d)" % self.params
        if self.op == "Zero":
            return "Zero(%d)" % self.dim
        if self.op == "IntCst":
            return "IntCst(%d)" % self.params
        if self.op == "RatCst":
            return "RatCst(%d,%d)" % self.params
        inner = ",".join([repr(a) for a in self.args] +
                         [("[%s]" % ",".join(map(str, p)) if isinstance(p, tuple) else repr(p)) for p in self.params])
        return "%s(%s)" % (self.op, inner)

    # ---- python operators (used by the parser and by autodiff) ---------------------------
    def __add__(self, o):
        return Add(self, as_node(o))

    def __radd__(self, o):
        return Add(as_node(o), self)

    def __sub__(self, o):
        return Subtract(self, as_node(o))

    def __rsub__(self, o):
        return Subtract(as_node(o), self)

    def __mul__(self, o):
        return Mult(self, as_node(o))

    def __rmul__(self, o):
        return Mult(as_node(o), self)

    def __truediv__(self, o):
        return Divide(self, as_node(o))

    def __rtruediv__(self, o):
        return Divide(as_node(o), self)

    def __neg__(self):
        return Minus(self)

    def __pow__(self, m):
        if isinstance(m, Node):
            return Powf(self, m)
        if m == 2:
            return Square(self)
        if isinstance(m, int):
            return Pow(self, m)
        if m == 0.5:
            return Sqrt(self)
        if m == -0.5:
            return Rsqrt(self)
        return Powf(self, FloatCst(float(m)))

    def __or__(self, o):
        return Scalprod(self, as_node(o))

    def __lt__(self, o):
        return LessThan(self, as_node(o))

    def __gt__(self, o):
        return LessThan(as_node(o), self)

    def __le__(self, o):
        return LessOrEqual(self, as_node(o))

    def __ge__(self, o):
        return LessOrEqual(as_node(o), self)

    __hash__ = object.__hash__

    def __eq__(self, o):  # identity (nodes are interned)
        return self is o

    def __ne__(self, o):
        return self is not o


def as_node(x) -> Node:
    if isinstance(x, Node):
        return x
    if isinstance(x, bool):
        raise FormulaError("invalid operand type bool")
    if isinstance(x, int):
        return IntCst(x)
    if isinstance(x, float):
        if x == int(x) and abs(x) < 2 ** 31:
            return IntCst(int(x))
        return FloatCst(x)
    raise FormulaError("invalid operand type : " + str(type(x)))


# ------------------------------------------------------------------------------------------
# leaves
# ------------------------------------------------------------------------------------------
def Var(ind: int, dim: int, cat: int) -> Node:
    """``Var(ind, dim, cat)``: cat 0 = i-indexed, 1 = j-indexed, 2 = parameter
    (``keopscore/keopscore/formulas/variables/Var.py:10-39``)."""
    if cat not in (0, 1, 2):
        raise FormulaError("Var category must be 0, 1 or 2")
    if dim < 1:
        raise FormulaError("Var dimension must be positive")
    return Node("Var", (), (int(ind), int(dim), int(cat)), dim)


def Zero(dim: int) -> Node:
    return Node("Zero", (), (), dim)


def IntCst(v: int) -> Node:
    return Node("IntCst", (), (int(v),), 1)


def RatCst(p: int, q: int) -> Node:
    if q == 0:
        raise FormulaError("division by zero")
    f = Fraction(int(p), int(q))
    if f.denominator == 1:
        return IntCst(f.numerator)
    return Node("RatCst", (), (f.numerator, f.denominator), 1)


def IntInv(v: int) -> Node:
    return RatCst(1, v)


def FloatCst(v: float) -> Node:
    return Node("FloatCst", (), (float(v),), 1)


def _const(fr: Fraction) -> Node:
    return RatCst(fr.numerator, fr.denominator)


# ------------------------------------------------------------------------------------------
# broadcasting helpers
# ------------------------------------------------------------------------------------------
def _bdim(*args: Node) -> int:
    dims = {a.dim for a in args}
    if len(dims) > 2 or (len(dims) == 2 and min(dims) != 1):
        raise FormulaError(
            "dimensions are not compatible for element-wise operation: %s" % sorted(dims)
        )
    return max(dims)


def Broadcast(a: Node, dim: int) -> Node:
    if a.dim == dim:
        return a
    if a.dim == 1:
        return SumT(a, dim)
    raise FormulaError("dimensions are not compatible for Broadcast operation")


# ------------------------------------------------------------------------------------------
# arithmetic
# ------------------------------------------------------------------------------------------
def Add(a: Node, b: Node) -> Node:
    d = _bdim(a, b)
    if a.is_zero:
        return Broadcast(b, d)
    if b.is_zero:
        return Broadcast(a, d)
    ca, cb = a.const_value(), b.const_value()
    if ca is not None and cb is not None:
        return _const(ca + cb)
    if a is b:
        return Mult(IntCst(2), a)
    return Node("Add", (a, b), (), d)


def Subtract(a: Node, b: Node) -> Node:
    d = _bdim(a, b)
    if a.is_zero:
        return Minus(Broadcast(b, d))
    if b.is_zero:
        return Broadcast(a, d)
    if a is b:
        return Zero(d)
    ca, cb = a.const_value(), b.const_value()
    if ca is not None and cb is not None:
        return _const(ca - cb)
    return Node("Subtract", (a, b), (), d)


def Minus(a: Node) -> Node:
    if a.is_zero:
        return a
    c = a.const_value()
    if c is not None:
        return _const(-c)
    if a.op == "Minus":
        return a.args[0]
    return Node("Minus", (a,), (), a.dim)


def Mult(a: Node, b: Node) -> Node:
    d = _bdim(a, b)
    if a.is_zero or b.is_zero:
        return Zero(d)
    ca, cb = a.const_value(), b.const_value()
    if ca is not None and cb is not None:
        return _const(ca * cb)
    if ca is not None:
        if ca == 1:
            return b
        if ca == -1:
            return Minus(b)
    if cb is not None:
        if cb == 1:
            return a
        if cb == -1:
            return Minus(a)
    if a is b:
        return Square(a)
    return Node("Mult", (a, b), (), d)


def Divide(a: Node, b: Node) -> Node:
    d = _bdim(a, b)
    if b.is_zero:
        raise FormulaError("division by zero")
    if a.is_zero:
        return Zero(d)
    cb = b.const_value()
    if cb is not None:
        # f / n -> (1/n) * f   (reference: ``maths/Divide.py:59-62``)
        return Mult(_const(1 / cb), a)
    ca = a.const_value()
    if ca is not None and ca == 1:
        return Inv(b)
    return Node("Divide", (a, b), (), d)


def _unary(op: str):
    def make(a: Node) -> Node:
        return Node(op, (a,), (), a.dim)

    make.__name__ = op
    return make


def Square(a: Node) -> Node:
    # constants are NOT folded here: the reference folds rational constants only in Add / Subtract / Mult / Minus /
    # Divide-by-constant (``maths/Square.py:27-31``, ``maths/Inv.py``), and a folded ``RatCst`` is emitted float-rounded in
    # every dtype (``variables/RatCst.py:22-24``) -- folding more than the reference does would change fp64 results
    # (``2 / (1 + Square(2))`` is an exact 0.4 there).  The scalar DAG folds the value itself, without the rounding.
    if a.is_zero:
        return a
    return Node("Square", (a,), (), a.dim)


def Inv(a: Node) -> Node:
    if a.is_zero:
        raise FormulaError("division by zero")
    return Node("Inv", (a,), (), a.dim)


Exp = _unary("Exp")
Log = _unary("Log")
Sin = _unary("Sin")
Cos = _unary("Cos")
Acos = _unary("Acos")
Asin = _unary("Asin")
Atan = _unary("Atan")
Floor = _unary("Floor")
XLogX = _unary("XLogX")
SinXDivX = _unary("SinXDivX")


def _unary_zero_fixed(op: str):
    def make(a: Node) -> Node:
        if a.is_zero:
            return a
        return Node(op, (a,), (), a.dim)

    make.__name__ = op
    return make


Sqrt = _unary_zero_fixed("Sqrt")
Rsqrt = _unary_zero_fixed("Rsqrt")  # reference guards x == 0 -> 0 (``math_functions.py:104-108``)
Abs = _unary_zero_fixed("Abs")
Sign = _unary_zero_fixed("Sign")
ReLU = _unary_zero_fixed("ReLU")
Step = _unary("Step")


def Pow(a: Node, m: int) -> Node:
    m = int(m)
    if m == 1:
        return a
    if m == 2:
        return Square(a)
    if m == 0:
        return Broadcast(IntCst(1), a.dim)
    return Node("Pow", (a,), (m,), a.dim)


def Powf(a: Node, b: Node) -> Node:
    return Node("Powf", (a, as_node(b)), (), _bdim(a, as_node(b)))


def Round(a: Node, d: int = 0) -> Node:
    return Node("Round", (a,), (int(d),), a.dim)


def Atan2(a: Node, b: Node) -> Node:
    return Node("Atan2", (a, b), (), _bdim(a, b))


def _cmp(op: str):
    def make(a: Node, b: Node) -> Node:
        return Node(op, (a, b), (), _bdim(a, b))

    make.__name__ = op
    return make


LessThan = _cmp("LessThan")
LessOrEqual = _cmp("LessOrEqual")
Equal = _cmp("Equal")
NotEqual = _cmp("NotEqual")


def IfElse(c: Node, a: Node, b: Node) -> Node:
    """``IfElse(f, a, b) = a if f >= 0 else b`` (``maths/IfElse.py``)."""
    d = _bdim(c, a, b)
    if a.is_zero and b.is_zero:
        return Zero(d)
    return Node("IfElse", (c, a, b), (), d)


def Clamp(x: Node, a: Node, b: Node) -> Node:
    return Node("Clamp", (x, a, b), (), _bdim(x, a, b))


def ClampInt(x: Node, a: int, b: int) -> Node:
    return Node("ClampInt", (x,), (int(a), int(b)), x.dim)


def DiffClampInt(x: Node, a: int, b: int) -> Node:
    return Node("DiffClampInt", (x,), (int(a), int(b)), x.dim)


def Mod(x: Node, n: Node, d: Node) -> Node:
    return Node("Mod", (x, n, d), (), _bdim(x, n, d))


# ------------------------------------------------------------------------------------------
# vector <-> scalar
# ------------------------------------------------------------------------------------------
def Sum(a: Node) -> Node:
    if a.dim == 1:
        return a
    if a.is_zero:
        return Zero(1)
    if a.op == "SumT":
        return Mult(IntCst(a.dim), a.args[0])
    return Node("Sum", (a,), (), 1)


def SumT(a: Node, dim: int) -> Node:
    if a.dim != 1:
        raise FormulaError("SumT expects a scalar formula")
    if dim == 1:
        return a
    if a.is_zero:
        return Zero(dim)
    return Node("SumT", (a,), (int(dim),), dim)


def Scalprod(a: Node, b: Node) -> Node:
    if a.dim != b.dim:
        raise FormulaError("Scalprod: dimensions must be the same")
    if a.is_zero or b.is_zero:
        return Zero(1)
    if a.dim == 1:
        return Mult(a, b)
    return Node("Scalprod", (a, b), (), 1)


def Max(a: Node) -> Node:
    return a if a.dim == 1 else Node("Max", (a,), (), 1)


def Min(a: Node) -> Node:
    return a if a.dim == 1 else Node("Min", (a,), (), 1)


def ArgMax(a: Node) -> Node:
    return Node("ArgMax", (a,), (), 1)


def ArgMin(a: Node) -> Node:
    return Node("ArgMin", (a,), (), 1)


def SqNorm2(a: Node) -> Node:
    return Sum(Square(a))


def Norm2(a: Node) -> Node:
    return Sqrt(Scalprod(a, a))


def Normalize(a: Node) -> Node:
    return Mult(Rsqrt(SqNorm2(a)), a)


def SqDist(a: Node, b: Node) -> Node:
    return SqNorm2(Subtract(a, b))


def WeightedSqNorm(A: Node, x: Node) -> Node:
    """``SqNormIso`` / ``SqNormDiag`` / ``SymSqNorm`` depending on ``A.dim``
    (``maths/WeightedSqNorm.py``)."""
    if A.dim == 1:
        return Mult(A, SqNorm2(x))
    if A.dim == x.dim:
        return Sum(Mult(A, Square(x)))
    if A.dim == x.dim * x.dim:
        return Scalprod(x, MatVecMult(A, x))
    raise FormulaError("WeightedSqNorm: incompatible dimensions")


def WeightedSqDist(A: Node, x: Node, y: Node) -> Node:
    return WeightedSqNorm(A, Subtract(x, y))


# ------------------------------------------------------------------------------------------
# shape ops
# ------------------------------------------------------------------------------------------
def Concat(a: Node, b: Node) -> Node:
    if a.is_zero and b.is_zero:
        return Zero(a.dim + b.dim)
    return Node("Concat", (a, b), (), a.dim + b.dim)


def Extract(a: Node, start: int, dim: int) -> Node:
    start, dim = int(start), int(dim)
    if start < 0 or start + dim > a.dim:
        raise FormulaError("Index out of bound in Extract")
    if a.is_zero:
        return Zero(dim)
    if start == 0 and dim == a.dim:
        return a
    return Node("Extract", (a,), (start, dim), dim)


def ExtractT(a: Node, start: int, dim: int) -> Node:
    start, dim = int(start), int(dim)
    if start < 0 or start + a.dim > dim:
        raise FormulaError("Index out of bound in ExtractT")
    if a.is_zero:
        return Zero(dim)
    if start == 0 and dim == a.dim:
        return a
    return Node("ExtractT", (a,), (start, dim), dim)


def Elem(a: Node, m: int) -> Node:
    m = int(m)
    if m < 0 or m >= a.dim:
        raise FormulaError("Index out of bound in Elem")
    return Extract(a, m, 1)


def ElemT(a: Node, n: int, m: int) -> Node:
    if a.dim != 1:
        raise FormulaError("Input of ElemT should be a scalar")
    return ExtractT(a, int(m), int(n))


def OneHot(a: Node, dim: int) -> Node:
    if a.dim != 1:
        raise FormulaError("OneHot operation is only valid for scalar input")
    return Node("OneHot", (a,), (int(dim),), int(dim))


def MatVecMult(A: Node, b: Node) -> Node:
    if A.dim % b.dim != 0:
        raise FormulaError("Dimensions of A and B are not compatible for matrix-vector product")
    return Node("MatVecMult", (A, b), (), A.dim // b.dim)


def VecMatMult(b: Node, A: Node) -> Node:
    if A.dim % b.dim != 0:
        raise FormulaError("Dimensions of A and B are not compatible for vector-matrix product")
    return Node("VecMatMult", (b, A), (), A.dim // b.dim)


def TensorProd(a: Node, b: Node) -> Node:
    if a.is_zero or b.is_zero:
        return Zero(a.dim * b.dim)
    return Node("TensorProd", (a, b), (), a.dim * b.dim)


# ------------------------------------------------------------------------------------------
# tensor contraction on the internal dimension (``keopscore/keopscore/formulas/maths/TensorDot.py``, ``Kron.py``)
# ------------------------------------------------------------------------------------------
def _row_major_strides(shape) -> List[int]:
    out, s = [], 1
    for n in reversed(shape):
        out.append(s)
        s *= n
    return out[::-1]


def _argsort(seq) -> List[int]:
    return sorted(range(len(seq)), key=seq.__getitem__)


class TensorDotLayout:
    """Index bookkeeping of ``TensorDot(a, b, dimsfa, dimsfb, contfa, contfb, permute)``.

    ``a`` / ``b`` are row-major tensors of shapes ``dimsfa`` / ``dimsfb`` stored flat; axes ``contfa[c]`` of ``a`` and
    ``contfb[c]`` of ``b`` are contracted pairwise.  The *natural* result has the kept axes of ``a`` (ascending) followed
    by the kept axes of ``b`` (as ``numpy.tensordot``); natural axis ``t`` is stored at axis ``permute[t]`` of the
    row-major output (``TensorDot.py:88-103,160-172``: the output strides are those of the shape ``natural`` scattered
    by ``permute``).  ``terms()`` lists, for every output component in natural row-major order, the ``(out, ia, ib)``
    flat indices of its products in the order the reference's loop nest adds them (kept axes outside, contracted
    axes inside in ``contfa`` order, ``TensorDot.py:155-187``).
    """

    def __init__(self, dimsfa, dimsfb, contfa, contfb, permute=None):
        self.dimsfa, self.dimsfb = [int(d) for d in dimsfa], [int(d) for d in dimsfb]
        self.contfa, self.contfb = [int(c) for c in contfa], [int(c) for c in contfb]
        if len(self.contfa) != len(self.contfb) or len(set(self.contfa)) != len(self.contfa) \
                or len(set(self.contfb)) != len(self.contfb) \
                or any(not 0 <= c < len(self.dimsfa) for c in self.contfa) \
                or any(not 0 <= c < len(self.dimsfb) for c in self.contfb) \
                or [self.dimsfa[c] for c in self.contfa] != [self.dimsfb[c] for c in self.contfb]:
            raise FormulaError("TensorDot: contracted dimensions of the two operands do not match")
        self.keep_a = [k for k in range(len(self.dimsfa)) if k not in self.contfa]
        self.keep_b = [k for k in range(len(self.dimsfb)) if k not in self.contfb]
        self.natural = [self.dimsfa[k] for k in self.keep_a] + [self.dimsfb[k] for k in self.keep_b]
        nk = len(self.natural)
        self.permute = list(range(nk)) if permute is None else [int(p) for p in permute]
        if sorted(self.permute) != list(range(nk)):
            raise FormulaError("TensorDot: permute must be a permutation of the %d kept axes" % nk)
        self.out_shape = [0] * nk
        for t, p in enumerate(self.permute):
            self.out_shape[p] = self.natural[t]
        self.cont_shape = [self.dimsfa[c] for c in self.contfa]
        self.dim = 1
        for n in self.natural:
            self.dim *= n

    def params(self) -> tuple:
        return (tuple(self.dimsfa), tuple(self.dimsfb), tuple(self.contfa), tuple(self.contfb), tuple(self.permute))

    def terms(self):
        import itertools
        sa, sb, so = _row_major_strides(self.dimsfa), _row_major_strides(self.dimsfb), _row_major_strides(self.out_shape)
        na = len(self.keep_a)
        groups = []
        for kept in itertools.product(*[range(n) for n in self.natural]):
            o = sum(kept[t] * so[self.permute[t]] for t in range(len(kept)))
            a0 = sum(kept[t] * sa[ax] for t, ax in enumerate(self.keep_a))
            b0 = sum(kept[na + t] * sb[ax] for t, ax in enumerate(self.keep_b))
            prods = []
            for cont in itertools.product(*[range(n) for n in self.cont_shape]):
                prods.append((a0 + sum(cont[c] * sa[ax] for c, ax in enumerate(self.contfa)),
                              b0 + sum(cont[c] * sb[ax] for c, ax in enumerate(self.contfb))))
            groups.append((o, prods))
        return groups

    def adjoint(self, wrt: int) -> tuple:
        """Parameters ``(dims_g, dims_other, cont_g, cont_other, permute)`` of the contraction of the incoming gradient
        (shape ``out_shape``) with the *other* operand that yields the gradient w.r.t. operand ``wrt`` in that operand's
        own layout (``TensorDot.py:105-139,189-216``)."""
        na = len(self.keep_a)
        if wrt == 0:
            cont_g = [self.permute[na + t] for t in range(len(self.keep_b))]
            back = [self.keep_a[t] for t in _argsort(self.permute[:na])] + \
                   [self.contfa[c] for c in _argsort(self.contfb)]
            return (self.out_shape, self.dimsfb, cont_g, self.keep_b, back)
        cont_g = [self.permute[t] for t in range(na)]
        back = [self.keep_b[t] for t in _argsort(self.permute[na:])] + \
               [self.contfb[c] for c in _argsort(self.contfa)]
        return (self.out_shape, self.dimsfa, cont_g, self.keep_a, back)


def _prod(seq) -> int:
    out = 1
    for n in seq:
        out *= int(n)
    return out


def TensorDot(a: Node, b: Node, dimsfa, dimsfb, contfa, contfb, permute=None) -> Node:
    lay = TensorDotLayout(dimsfa, dimsfb, contfa, contfb, permute)
    if a.dim != _prod(lay.dimsfa) or b.dim != _prod(lay.dimsfb):
        raise FormulaError("TensorDot: operand dimensions %d, %d do not match the shapes %s, %s"
                           % (a.dim, b.dim, lay.dimsfa, lay.dimsfb))
    if a.is_zero or b.is_zero:
        return Zero(lay.dim)
    return Node("TensorDot", (a, b), lay.params(), lay.dim)


def Kron(a: Node, b: Node, dimsfa, dimsfb) -> Node:
    """``numpy.kron`` of two tensors with the same number of axes = an outer product whose axes are interleaved
    (``maths/Kron.py:17-31``)."""
    if len(dimsfa) != len(dimsfb):
        raise FormulaError("Kron: the two operands must have the same number of axes")
    n = len(dimsfa)
    return TensorDot(a, b, dimsfa, dimsfb, (), (), list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2)))


# ------------------------------------------------------------------------------------------
# soft dynamic time warping between two vectors (``keopscore/keopscore/formulas/maths/SoftDTW_SqDist.py``)
# ------------------------------------------------------------------------------------------
def SoftDTW_SqDist(x: Node, y: Node, gamma: Node) -> Node:
    """Soft-DTW discrepancy (soft-min of smoothness ``gamma`` over all monotone alignments of the entries of ``x`` and
    ``y``, squared difference as the ground cost) -- a scalar (``SoftDTW_SqDist.py:19-93``)."""
    if gamma.dim != 1:
        raise FormulaError("input gamma should be scalar")
    return Node("SoftDTW_SqDist", (x, y, gamma), (), 1)


def GradSoftDTW_SqDist(x: Node, y: Node, gamma: Node) -> Node:
    """Gradient of :func:`SoftDTW_SqDist` with respect to ``x`` and ``y``, concatenated (``SoftDTW_SqDist.py:109-216``)."""
    if gamma.dim != 1:
        raise FormulaError("input gamma should be scalar")
    return Node("GradSoftDTW_SqDist", (x, y, gamma), (), x.dim + y.dim)


def BSpline(knots: Node, x: Node, order: int) -> Node:
    """The ``knots.dim - order - 1`` B-spline basis functions of the given order on the knot vector ``knots``, sampled at the
    scalar ``x`` (de Boor's recursion, ``keopscore/keopscore/formulas/maths/BSpline.py:19-126``)."""
    order = int(order)
    if order < 0:
        raise FormulaError("The order of a BSpline must be >= 0, but received %d." % order)
    if x.dim != 1:
        raise FormulaError("BSplines must be sampled at scalar values, but received a vector of size %d." % x.dim)
    if knots.dim < order + 2:
        raise FormulaError("BSplines of order %d require at least %d knots, but only received a vector of size %d."
                           % (order, order + 2, knots.dim))
    return Node("BSpline", (knots, x), (order,), knots.dim - order - 1)


# ------------------------------------------------------------------------------------------
# complex arithmetic on interleaved (re, im) pairs (``keopscore/keopscore/formulas/complex/*.py``)
# ------------------------------------------------------------------------------------------
def _even(f: Node, what: str) -> None:
    if f.dim % 2 != 0:
        raise FormulaError("Dimension of F must be even (%s of a vector of size %d)" % (what, f.dim))


def _complex_dims(name: str, *args: Node) -> int:
    """Broadcast rule of the reference's ``VectorizedComplexScalarOp`` (``formulas/VectorizedComplexScalarOp.py:11-21``):
    all operands have the same even size, or some are one complex number (size 2)."""
    dims = set(a.dim for a in args)
    if max(dims) % 2 != 0 or len(dims) > 2 or (len(dims) == 2 and min(dims) != 2):
        raise FormulaError("dimensions are not compatible for VectorizedComplexScalarOp (%s)" % name)
    return max(dims)


def ComplexReal(f: Node) -> Node:
    _even(f, "ComplexReal")
    return Zero(f.dim // 2) if f.is_zero else Node("ComplexReal", (f,), (), f.dim // 2)


def ComplexImag(f: Node) -> Node:
    _even(f, "ComplexImag")
    return Zero(f.dim // 2) if f.is_zero else Node("ComplexImag", (f,), (), f.dim // 2)


def Real2Complex(f: Node) -> Node:
    return Zero(2 * f.dim) if f.is_zero else Node("Real2Complex", (f,), (), 2 * f.dim)


def Imag2Complex(f: Node) -> Node:
    return Zero(2 * f.dim) if f.is_zero else Node("Imag2Complex", (f,), (), 2 * f.dim)


def Conj(f: Node) -> Node:
    _complex_dims("Conj", f)
    return f if f.is_zero else Node("Conj", (f,), (), f.dim)


def ComplexSquareAbs(f: Node) -> Node:
    _even(f, "ComplexSquareAbs")
    return Zero(f.dim // 2) if f.is_zero else Node("ComplexSquareAbs", (f,), (), f.dim // 2)


def ComplexAbs(f: Node) -> Node:
    return Sqrt(ComplexSquareAbs(f))


def ComplexAngle(f: Node) -> Node:
    return Atan2(ComplexImag(f), ComplexReal(f))


def ComplexMult(f: Node, g: Node) -> Node:
    dim = _complex_dims("ComplexMult", f, g)
    if f.is_zero or g.is_zero:
        return Zero(dim)
    return Node("ComplexMult", (f, g), (), dim)


def ComplexDivide(f: Node, g: Node) -> Node:
    return ComplexMult(Real2Complex(Inv(ComplexSquareAbs(g))), ComplexMult(f, Conj(g)))


def ComplexAdd(f: Node, g: Node) -> Node:
    return Node("ComplexAdd", (f, g), (), _complex_dims("ComplexAdd", f, g))


def ComplexSubtract(f: Node, g: Node) -> Node:
    return Node("ComplexSubtract", (f, g), (), _complex_dims("ComplexSubtract", f, g))


def ComplexRealScal(f: Node, g: Node) -> Node:
    """real scalar times complex vector, in either argument order (``complex/ComplexRealScal.py:54-55``)"""
    if f.dim != 1:
        f, g = g, f
    if f.dim != 1:
        raise FormulaError("Dimension of F must be 1")
    _even(g, "ComplexRealScal")
    if f.is_zero or g.is_zero:
        return Zero(g.dim)
    return Node("ComplexRealScal", (f, g), (), g.dim)


def ComplexExp(f: Node) -> Node:
    return Node("ComplexExp", (f,), (), _complex_dims("ComplexExp", f))


def ComplexExp1j(f: Node) -> Node:
    return Node("ComplexExp1j", (f,), (), 2 * f.dim)


def ComplexSum(f: Node) -> Node:
    _even(f, "ComplexSum")
    return Zero(2) if f.is_zero else Node("ComplexSum", (f,), (), 2)


def ComplexSumT(f: Node, dim: int) -> Node:
    if f.dim != 2:
        raise FormulaError("Dimension of F must be 2")
    dim = int(dim)
    if dim % 2 != 0:
        raise FormulaError("ComplexSumT: the output dimension must be even")
    return Zero(dim) if f.is_zero else Node("ComplexSumT", (f,), (dim,), dim)


# names visible to the formula parser -------------------------------------------------------
PUBLIC_OPS = {
    name: obj
    for name, obj in list(globals().items())
    if callable(obj) and name[0].isupper() and name not in ("Node", "FormulaError", "Fraction")
}
for _hidden in ("Broadcast", "Dict", "Iterable", "List", "Optional", "Tuple", "TensorDotLayout"):
    PUBLIC_OPS.pop(_hidden, None)
