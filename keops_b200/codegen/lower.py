"""Lowering of the vector formula IR to the scalar DAG (one signed scalar per component).

Summation order follows the reference's generated code so that value orderings (ArgMin / ArgKMin)
agree: ``Sum`` and ``Scalprod`` accumulate left to right starting from the first term
(``keopscore/keopscore/formulas/maths/Sum.py:20-24``), and products feeding an
addition are contracted to FMAs by the emitter exactly as nvcc contracts the reference's code.
"""
from __future__ import annotations

import struct

from typing import Dict, List

from ..formulas.ir import Node, FormulaError
from .scalar import Graph, SV


class Lowerer:
    def __init__(self, graph: Graph):
        self.g = graph
        self.memo: Dict[int, List[SV]] = {}

    def __call__(self, node: Node) -> List[SV]:
        key = id(node)
        out = self.memo.get(key)
        if out is None:
            out = self._lower(node)
            if len(out) != node.dim:
                raise FormulaError("internal: lowering of %s produced %d components, expected %d"
                                   % (node.op, len(out), node.dim))
            self.memo[key] = out
        return out

    def _bc(self, vals: List[SV], dim: int, k: int) -> SV:
        return vals[k] if len(vals) > 1 else vals[0]

    def _lower(self, n: Node) -> List[SV]:
        g, op = self.g, n.op
        L = self

        if op == "Var":
            ind, dim, cat = n.params
            return [g.load(ind, k, cat) for k in range(dim)]
        if op == "Zero":
            return [g.const(0.0)] * n.dim
        if op == "IntCst":
            return [g.const(float(n.params[0]))]
        if op == "RatCst":
            # reference quirk: a rational constant is emitted as ``(float)(p/q)`` whatever the dtype
            # (keopscore/keopscore/formulas/variables/RatCst.py:22-24) -- and ``x / IntCst(n)`` is rewritten to
            # ``RatCst(1, n) * x`` (maths/Divide.py:52-53) -- so in float64 these constants carry float32 rounding.
            # Reproduced here: parity with the reference's outputs is the contract (golden case op2_extractT_cst_float64).
            return [g.const(struct.unpack("f", struct.pack("f", n.params[0] / n.params[1]))[0])]
        if op == "FloatCst":
            return [g.const(n.params[0])]

        if op in ("Add", "Subtract", "Mult", "Divide", "Powf", "Atan2", "LessThan", "LessOrEqual",
                  "Equal", "NotEqual"):
            a, b = L(n.args[0]), L(n.args[1])
            out = []
            for k in range(n.dim):
                x, y = self._bc(a, n.dim, k), self._bc(b, n.dim, k)
                if op == "Add":
                    out.append(g.add(x, y))
                elif op == "Subtract":
                    out.append(g.sub(x, y))
                elif op == "Mult":
                    out.append(g.mul(x, y))
                elif op == "Divide":
                    out.append(g.div(x, y))
                elif op == "Powf":
                    out.append(g.fn2("powf", x, y))
                elif op == "Atan2":
                    out.append(g.fn2("atan2", x, y))
                elif op == "LessThan":
                    out.append(g.fn2("lt", x, y))
                elif op == "LessOrEqual":
                    out.append(g.fn2("le", x, y))
                elif op == "Equal":
                    out.append(g.fn2("eq", x, y))
                else:
                    out.append(g.fn2("ne", x, y))
            return out

        if op in _UNARY:
            (a,) = [L(x) for x in n.args]
            return [_UNARY[op](g, x, n) for x in a]

        if op == "IfElse":
            c, a, b = [L(x) for x in n.args]
            return [g.select_ge0(self._bc(c, n.dim, k), self._bc(a, n.dim, k), self._bc(b, n.dim, k))
                    for k in range(n.dim)]
        if op == "Clamp":
            x, lo, hi = [L(v) for v in n.args]
            return [g.fn3("clamp", self._bc(x, n.dim, k), self._bc(lo, n.dim, k), self._bc(hi, n.dim, k))
                    for k in range(n.dim)]
        if op == "Mod":
            x, m, d = [L(v) for v in n.args]
            return [g.fn3("mod", self._bc(x, n.dim, k), self._bc(m, n.dim, k), self._bc(d, n.dim, k))
                    for k in range(n.dim)]

        if op == "Sum":
            a = L(n.args[0])
            acc = a[0]
            for x in a[1:]:
                acc = g.add(acc, x)
            return [acc]
        if op == "SumT":
            a = L(n.args[0])
            return [a[0]] * n.dim
        if op == "Scalprod":
            a, b = L(n.args[0]), L(n.args[1])
            acc = g.mul(a[0], b[0])
            for x, y in zip(a[1:], b[1:]):
                acc = g.add(acc, g.mul(x, y))
            return [acc]
        if op in ("Max", "Min"):
            a = L(n.args[0])
            acc = a[0]
            for x in a[1:]:
                acc = g.fn2("fmax" if op == "Max" else "fmin", acc, x)
            return [acc]
        if op in ("ArgMax", "ArgMin"):
            a = L(n.args[0])
            best, arg = a[0], g.const(0.0)
            for k, x in enumerate(a[1:], start=1):
                # strict comparison: first extremum wins (reference ``maths/ArgMax.py`` / ``ArgMin.py``)
                better = g.fn2("lt", best, x) if op == "ArgMax" else g.fn2("lt", x, best)
                # better is 1.0 / 0.0 ; select via (better - 0.5) >= 0
                sel = g.sub(better, g.const(0.5))
                arg = g.select_ge0(sel, g.const(float(k)), arg)
                best = g.select_ge0(sel, x, best)
            return [arg]
        if op == "OneHot":
            a = L(n.args[0])[0]
            r = g.fn1("round0", a)
            return [g.fn2("eq", r, g.const(float(k))) for k in range(n.dim)]
        if op == "Concat":
            return L(n.args[0]) + L(n.args[1])
        if op == "Extract":
            start, dim = n.params
            return L(n.args[0])[start:start + dim]
        if op == "ExtractT":
            start, dim = n.params
            a = L(n.args[0])
            z = g.const(0.0)
            return [z] * start + a + [z] * (dim - start - len(a))
        if op == "MatVecMult":
            A, b = L(n.args[0]), L(n.args[1])
            m = len(b)
            rows = len(A) // m
            out = []
            for r in range(rows):
                acc = g.mul(A[r * m], b[0])
                for c in range(1, m):
                    acc = g.add(acc, g.mul(A[r * m + c], b[c]))
                out.append(acc)
            return out
        if op == "VecMatMult":
            b, A = L(n.args[0]), L(n.args[1])
            m = len(b)
            cols = len(A) // m
            out = []
            for c in range(cols):
                acc = g.mul(b[0], A[c])
                for r in range(1, m):
                    acc = g.add(acc, g.mul(b[r], A[r * cols + c]))
                out.append(acc)
            return out
        if op == "TensorProd":
            a, b = L(n.args[0]), L(n.args[1])
            return [g.mul(x, y) for x in a for y in b]
        if op.startswith("Complex") or op in ("Real2Complex", "Imag2Complex", "Conj"):
            return self._lower_complex(n)
        if op == "BSpline":
            # de Boor's recursion, in place on knots.dim - 1 values; the first knots.dim - order - 1 are the result
            # (maths/BSpline.py:44-95).  omega(i, k) = (x - t_i) / (t_{i+k} - t_i), 0 on an empty knot span.
            t, x = L(n.args[0]), L(n.args[1])[0]
            (order,) = n.params
            half, zero, one = g.const(0.5), g.const(0.0), g.const(1.0)

            def omega(i, k):
                nonempty = g.sub(g.fn2("lt", t[i], t[i + k]), half)
                return g.select_ge0(nonempty, g.div(g.sub(x, t[i]), g.sub(t[i + k], t[i])), zero)

            out = [g.mul(g.fn2("le", t[j], x), g.fn2("lt", x, t[j + 1])) for j in range(len(t) - 1)]
            for k in range(1, order + 1):
                w1 = omega(0, k)
                for i in range(len(out) - k):
                    w2 = omega(i + 1, k)
                    out[i] = g.add(g.mul(w1, out[i]), g.mul(g.sub(one, w2), out[i + 1]))
                    w1 = w2
            return out[:n.dim]
        if op in ("SoftDTW_SqDist", "GradSoftDTW_SqDist"):
            x, y, gam = L(n.args[0]), L(n.args[1]), L(n.args[2])[0]
            r = _softdtw_table(g, x, y, gam)
            if op == "SoftDTW_SqDist":
                return [r[len(x) - 1][len(y) - 1]]
            return _softdtw_gradient(g, x, y, gam, r)
        if op == "TensorDot":
            from ..formulas.ir import TensorDotLayout
            a, b = L(n.args[0]), L(n.args[1])
            out = [None] * n.dim
            for o, prods in TensorDotLayout(*n.params).terms():
                (ia, ib), rest = prods[0], prods[1:]
                acc = g.mul(a[ia], b[ib])
                for ia, ib in rest:
                    acc = g.add(acc, g.mul(a[ia], b[ib]))
                out[o] = acc
            return out
        raise NotImplementedError("operation %s is not supported by the sm_100a code generator" % op)


def _lower_complex(self, n: Node) -> List[SV]:
    """Complex arithmetic on interleaved (re, im) components (``keopscore/keopscore/formulas/complex/*.py``); a size-2 operand
    of a binary operation is one complex number broadcast over the other operand's pairs."""
    g, op = self.g, n.op
    a = self(n.args[0])
    zero = g.const(0.0)
    pairs = lambda v: [(v[k], v[k + 1]) for k in range(0, len(v), 2)]
    flat = lambda ps: [c for p in ps for c in p]
    if op == "ComplexReal":
        return a[0::2]
    if op == "ComplexImag":
        return a[1::2]
    if op == "Real2Complex":
        return flat((x, zero) for x in a)
    if op == "Imag2Complex":
        return flat((zero, x) for x in a)
    if op == "Conj":
        return flat((re, -im) for re, im in pairs(a))
    if op == "ComplexSquareAbs":
        return [g.add(g.mul(re, re), g.mul(im, im)) for re, im in pairs(a)]
    if op == "ComplexExp":
        out = []
        for re, im in pairs(a):
            r = g.exp(re)
            out += [g.mul(r, g.fn1("cos", im, parity=1)), g.mul(r, g.fn1("sin", im, parity=-1))]
        return out
    if op == "ComplexExp1j":
        return flat((g.fn1("cos", x, parity=1), g.fn1("sin", x, parity=-1)) for x in a)
    if op == "ComplexSum":
        ps = pairs(a)
        re, im = ps[0]
        for r2, i2 in ps[1:]:
            re, im = g.add(re, r2), g.add(im, i2)
        return [re, im]
    if op == "ComplexSumT":
        return list(a) * (n.dim // 2)
    b = self(n.args[1])
    if op == "ComplexRealScal":
        return [g.mul(a[0], x) for x in b]
    pa, pb = pairs(a), pairs(b)
    npair = n.dim // 2
    pa = pa * npair if len(pa) == 1 else pa
    pb = pb * npair if len(pb) == 1 else pb
    if op == "ComplexAdd":
        return flat((g.add(x[0], y[0]), g.add(x[1], y[1])) for x, y in zip(pa, pb))
    if op == "ComplexSubtract":
        return flat((g.sub(x[0], y[0]), g.sub(x[1], y[1])) for x, y in zip(pa, pb))
    if op == "ComplexMult":
        return flat((g.sub(g.mul(x[0], y[0]), g.mul(x[1], y[1])), g.add(g.mul(x[0], y[1]), g.mul(x[1], y[0])))
                    for x, y in zip(pa, pb))
    raise NotImplementedError("operation %s is not supported by the sm_100a code generator" % op)


Lowerer._lower_complex = _lower_complex


def _softdtw_table(g: Graph, x: List[SV], y: List[SV], gam: SV) -> List[List[SV]]:
    """``r[i][j]``: soft-DTW value of the prefixes ``x[:i+1]``, ``y[:j+1]`` -- the recursion, the operand order of the
    three-term soft-min and its shift by the hard minimum follow ``maths/SoftDTW_SqDist.py:33-93`` (which keeps one column
    in registers; the values are the same)."""
    n, m = len(x), len(y)

    def cost(i, j):
        return g.square(g.sub(x[i], y[j]))

    r = [[None] * m for _ in range(n)]
    r[0][0] = cost(0, 0)
    for i in range(1, n):
        r[i][0] = g.add(cost(i, 0), r[i - 1][0])
    for j in range(1, m):
        r[0][j] = g.add(cost(0, j), r[0][j - 1])
        for i in range(1, n):
            diag, left, up = r[i - 1][j - 1], r[i][j - 1], r[i - 1][j]
            lo = g.fn2("fmin", g.fn2("fmin", diag, left), up)
            s = g.add(g.add(g.exp(g.div(g.sub(lo, diag), gam)), g.exp(g.div(g.sub(lo, up), gam))),
                      g.exp(g.div(g.sub(lo, left), gam)))
            r[i][j] = g.add(cost(i, j), g.sub(lo, g.mul(gam, g.log(s))))
    return r


def _softdtw_gradient(g: Graph, x: List[SV], y: List[SV], gam: SV, r: List[List[SV]]) -> List[SV]:
    """Backward sweep over the alignment matrix: ``e[i][j]`` = d r[n-1][m-1] / d r[i][j]``, then the chain rule through the
    ground costs (``maths/SoftDTW_SqDist.py:159-211``); accumulation orders as there (columns from the last to the first,
    rows from the last to the first inside a column)."""
    n, m = len(x), len(y)
    two = g.const(2.0)

    def diff(i, j):
        return g.sub(x[i], y[j])

    def weight(i1, j1, i, j):
        # exp((r[i1][j1] - r[i][j] - cost(i1, j1)) / gamma): share of (i, j) in the soft-min that produced r[i1][j1]
        return g.exp(g.div(g.sub(g.sub(r[i1][j1], r[i][j]), g.square(diff(i1, j1))), gam))

    e = [[None] * m for _ in range(n)]
    gx, gy = [None] * n, [None] * m

    def emit(i, j):
        t = g.mul(g.mul(two, diff(i, j)), e[i][j])
        gx[i] = t if gx[i] is None else g.add(gx[i], t)
        gy[j] = -t if gy[j] is None else g.sub(gy[j], t)

    e[n - 1][m - 1] = g.const(1.0)
    emit(n - 1, m - 1)
    for i in range(n - 2, -1, -1):
        e[i][m - 1] = g.mul(weight(i + 1, m - 1, i, m - 1), e[i + 1][m - 1])
        emit(i, m - 1)
    for j in range(m - 2, -1, -1):
        e[n - 1][j] = g.mul(weight(n - 1, j + 1, n - 1, j), e[n - 1][j + 1])
        emit(n - 1, j)
        for i in range(n - 2, -1, -1):
            a, b, c = weight(i + 1, j, i, j), weight(i, j + 1, i, j), weight(i + 1, j + 1, i, j)
            e[i][j] = g.add(g.add(g.mul(a, e[i + 1][j]), g.mul(b, e[i][j + 1])), g.mul(c, e[i + 1][j + 1]))
            emit(i, j)
    return gx + gy


def _pow(g: Graph, x: SV, n: Node) -> SV:
    return g.powi(x, n.params[0])


def _round(g: Graph, x: SV, n: Node) -> SV:
    d = n.params[0]
    if d == 0:
        return g.fn1("round0", x, parity=-1)
    s = float(10 ** d)
    return g.mul(g.fn1("round0", g.mul(x, g.const(s)), parity=-1), g.const(1.0 / s))


def _clampint(g: Graph, x: SV, n: Node) -> SV:
    lo, hi = n.params
    return g.fn3("clamp", x, g.const(float(lo)), g.const(float(hi)))


def _diffclampint(g: Graph, x: SV, n: Node) -> SV:
    lo, hi = n.params
    return g.fn3("inrange", x, g.const(float(lo)), g.const(float(hi)))


_UNARY = {
    "Minus": lambda g, x, n: -x,
    "Square": lambda g, x, n: g.square(x),
    "Inv": lambda g, x, n: g.rcp(x),
    "Exp": lambda g, x, n: g.exp(x),
    "Log": lambda g, x, n: g.log(x),
    "Sqrt": lambda g, x, n: g.sqrt(x),
    "Rsqrt": lambda g, x, n: g.rsqrt(x),
    "Sin": lambda g, x, n: g.fn1("sin", x, parity=-1),
    "Cos": lambda g, x, n: g.fn1("cos", x, parity=1),
    "Asin": lambda g, x, n: g.fn1("asin", x, parity=-1),
    "Acos": lambda g, x, n: g.fn1("acos", x),
    "Atan": lambda g, x, n: g.fn1("atan", x, parity=-1),
    "Abs": lambda g, x, n: g.fn1("abs", x, parity=1),
    "Sign": lambda g, x, n: g.fn1("sign", x, parity=-1),
    "Step": lambda g, x, n: g.fn1("step", x),
    "ReLU": lambda g, x, n: g.fn1("relu", x),
    "Floor": lambda g, x, n: g.fn1("floor", x),
    "XLogX": lambda g, x, n: g.fn1("xlogx", x),
    "SinXDivX": lambda g, x, n: g.fn1("sinxdivx", x, parity=1),
    "Pow": _pow,
    "Round": _round,
    "ClampInt": _clampint,
    "DiffClampInt": _diffclampint,
}
