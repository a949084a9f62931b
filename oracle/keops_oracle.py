"""CPU oracle for the KeOps map-reduce hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product (``keops_b200/``) never does and has no CPU fallback.

It is a dense numpy restatement of what the reference computes for
``out_i = Reduce_j F(x_i, y_j, p)``:

* the formula string is ``eval``'d in a namespace of operation classes, like the reference's
  ``GetReduction`` (``keopscore/keopscore/formulas/GetReduction.py:19-39``);
* each operation evaluates itself on broadcastable ``(M, N, dim)`` arrays in the formula dtype with
  the reference CPU code's operation order -- e.g. ``Sum`` / ``Scalprod`` accumulate left to right,
  ``out = 0; out += arg[k]`` (``formulas/maths/Sum.py:20-24``), and no FMA contraction (the CPU
  backend is built without ``-march=native``, ``config/base_config.py:266-291``);
* each operation knows its adjoint ``difft(v, gradin)`` (the reference's ``DiffT``), so that
  ``Grad_WithSavedForward`` strings -- what ``GenredAutograd._backward`` hands to the engine
  (``pykeops/pykeops/torch/generic/generic_red.py:165-189``) -- can be evaluated too;
* reductions follow ``keopscore/keopscore/formulas/reductions/*.py`` and the scalar loop of
  ``CpuReduc`` (``keopscore/keopscore/mapreduce/cpu/CpuReduc.py:46-67``): ties in Min / ArgMin /
  KMin go to the smallest index; arg-type outputs are stored as floats; ``Max_SumShiftExp``
  returns the unnormalised pair ``(m, s)``.  Sums over ``j`` are accumulated in float64 and rounded
  once (the reference uses float32 block sums of ~sqrt(N) terms; both agree to ~1e-7 relative).

Parity pinning: ``tests/test_oracle_vs_golden.py`` checks this module against outputs of the real
reference (CPU backend, run in the build container by ``tests/golden/make_golden.py``) and against
the reference's one hard-coded known answer (``pykeops/pykeops/torch/test_install.py:5-25``).
"""
from __future__ import annotations

import re
from fractions import Fraction

import numpy as np

# ------------------------------------------------------------------------------------------
# operations
# ------------------------------------------------------------------------------------------


class Op:
    dim = 1
    children = ()

    # python operators, as in ``formulas/Operation.py:118-216``
    def __add__(s, o): return Add(s, _w(o))
    def __radd__(s, o): return Add(_w(o), s)
    def __sub__(s, o): return Subtract(s, _w(o))
    def __rsub__(s, o): return Subtract(_w(o), s)
    def __mul__(s, o): return Mult(s, _w(o))
    def __rmul__(s, o): return Mult(_w(o), s)
    def __truediv__(s, o): return Divide(s, _w(o))
    def __rtruediv__(s, o): return Divide(_w(o), s)
    def __neg__(s): return Minus(s)
    def __or__(s, o): return Scalprod(s, _w(o))
    def __lt__(s, o): return LessThan(s, _w(o))
    def __gt__(s, o): return LessThan(_w(o), s)
    def __le__(s, o): return LessOrEqual(s, _w(o))
    def __ge__(s, o): return LessOrEqual(_w(o), s)

    def __pow__(s, m):
        if isinstance(m, Op):
            return Powf(s, m)
        if m == 2:
            return Square(s)
        if isinstance(m, int):
            return Pow(s, m)
        if m == 0.5:
            return Sqrt(s)
        if m == -0.5:
            return Rsqrt(s)
        return Powf(s, Cst(float(m)))

    def vars(self):
        out = {}
        for c in self.children:
            out.update(c.vars())
        return out

    def iszero(self):
        return False


def _w(x):
    if isinstance(x, Op):
        return x
    if isinstance(x, (int, float)):
        return Cst(x)
    raise ValueError("invalid operand " + repr(x))


class Var(Op):
    def __init__(self, ind, dim, cat):
        self.ind, self.dim, self.cat = int(ind), int(dim), int(cat)

    def vars(self):
        return {self.ind: self}

    def ev(self, env):
        return env.var(self)

    def difft(self, v, g):
        same = (self.ind, self.dim, self.cat) == (v.ind, v.dim, v.cat)
        return g if same else Zero(v.dim)


class Zero(Op):
    def __init__(self, dim):
        self.dim = int(dim)

    def iszero(self):
        return True

    def ev(self, env):
        return np.zeros((1, 1, self.dim), env.dtype)

    def difft(self, v, g):
        return Zero(v.dim)


class Cst(Op):
    """IntCst / RatCst / float constant."""

    def __init__(self, val):
        self.val = val

    def ev(self, env):
        val = float(self.val)
        if isinstance(self.val, Fraction) and self.val.denominator != 1:
            # a rational constant is emitted as ``(float)(p/q)`` whatever the dtype
            # (keopscore/keopscore/formulas/variables/RatCst.py:22-24)
            val = float(np.float32(val))
        return np.full((1, 1, 1), val, env.dtype)

    def difft(self, v, g):
        return Zero(v.dim)


def IntCst(v): return Cst(int(v))
def RatCst(p, q): return Cst(Fraction(int(p), int(q)))
def IntInv(v): return Cst(Fraction(1, int(v)))


def _bd(*ops):
    return max(o.dim for o in ops)


def _red(g, dim):
    """adjoint of broadcasting dim -> g.dim"""
    if g.dim == dim:
        return g
    if dim == 1:
        return Sum(g)
    raise ValueError("cannot reduce gradient")


def _z(*xs):
    return any(x.iszero() for x in xs)


class _Elementwise(Op):
    def __init__(self, *children):
        self.children = children
        self.dim = _bd(*children)

    def ev(self, env):
        return self.f(env, *[c.ev(env) for c in self.children])


class _Add(_Elementwise):

    def f(self, env, a, b): return a + b
    def difft(self, v, g):
        a, b = self.children
        return a.difft(v, _red(g, a.dim)) + b.difft(v, _red(g, b.dim))          # maths/Add.py



def _rat(x):
    """exact value of an integer / rational constant node, else None"""
    return Fraction(x.val) if isinstance(x, Cst) and isinstance(x.val, (int, Fraction)) else None


def _mkcst(fr):
    """the reference's IntCst / RatCst constructors: 0 -> Zero, q == 1 -> IntCst (variables/RatCst.py:31-39, IntCst.py:30-34)"""
    fr = Fraction(fr)
    return Zero(1) if fr == 0 else Cst(int(fr) if fr.denominator == 1 else fr)


def Add(a, b):
    if a.iszero():
        return Broadcast(b, _bd(a, b))
    if b.iszero():
        return Broadcast(a, _bd(a, b))
    if _rat(a) is not None and _rat(b) is not None:                              # maths/Add.py:64-77: m + n, m + p/q, ...
        return _mkcst(_rat(a) + _rat(b))
    return _Add(a, b)

class _Subtract(_Elementwise):

    def f(self, env, a, b): return a - b
    def difft(self, v, g):
        a, b = self.children
        return a.difft(v, _red(g, a.dim)) - b.difft(v, _red(g, b.dim))          # maths/Subtract.py:26-33



def Subtract(a, b):
    if b.iszero():
        return Broadcast(a, _bd(a, b))
    if a.iszero():
        return Minus(Broadcast(b, _bd(a, b)))
    if _rat(a) is not None and _rat(b) is not None:                              # maths/Subtract.py:62-75
        return _mkcst(_rat(a) - _rat(b))
    return _Subtract(a, b)

class _Minus(_Elementwise):

    def f(self, env, a): return -a
    def difft(self, v, g):
        return -(self.children[0].difft(v, g))                                   # maths/Minus.py:25-27



def Minus(a):
    if a.iszero():
        return a
    if _rat(a) is not None:                                                      # maths/Minus.py:48-54: -(n), -(p/q)
        return _mkcst(-_rat(a))
    return _Minus(a)

class _Mult(_Elementwise):

    def f(self, env, a, b): return a * b
    def difft(self, v, g):
        a, b = self.children                                                     # maths/Mult.py:28-38
        return a.difft(v, _red(g * b, a.dim)) + b.difft(v, _red(g * a, b.dim))



def Mult(a, b):
    if _z(a, b):
        return Zero(_bd(a, b))
    if _rat(a) is not None and _rat(b) is not None:                              # maths/Mult.py:60-78: mn, m(p/q), (p/q)(m/n)
        return _mkcst(_rat(a) * _rat(b))
    if isinstance(b, Cst) and isinstance(b.val, int):                            # f*n -> n*f, integers only (maths/Mult.py:73-75)
        return Mult(b, a)
    if isinstance(b, _Mult) and _rat(b.children[0]) is not None:                 # maths/Mult.py:79-93
        c, g = b.children
        if _rat(a) is not None and (isinstance(a.val, int) or isinstance(c.val, Fraction)):
            return Mult(Mult(a, c), g)                                           # m*(n*g) -> (mn)*g,  m*(r*g) -> (m*r)*g
        return Mult(c, Mult(a, g))                                               # f*(n*g) -> n*(f*g)
    if isinstance(a, _Mult) and _rat(a.children[0]) is not None and not isinstance(b, Exp):
        return Mult(a.children[0], Mult(a.children[1], b))                       # (n*f)*g -> n*(f*g)  (maths/Mult.py:94-99)
    return _Mult(a, b)

class _Divide(_Elementwise):

    def f(self, env, a, b): return a / b
    def difft(self, v, g):
        a, b = self.children                                                     # maths/Divide.py:28-37
        return a.difft(v, _red(g / b, a.dim)) - b.difft(v, _red(g * a / Square(b), b.dim))



def Divide(a, b):
    if a.iszero():
        return Zero(_bd(a, b))
    if isinstance(b, Cst) and isinstance(b.val, (int, Fraction)) and b.val != 0:
        # division by an integer / rational constant becomes a product with its (float-rounded) inverse
        # (keopscore/keopscore/formulas/maths/Divide.py:52-55)
        return Cst(1 / Fraction(b.val)) * a
    return _Divide(a, b)

def Broadcast(a, dim):
    if a.dim == dim:
        return a
    if a.dim == 1:
        return SumT(a, dim)
    raise ValueError("dimensions are not compatible for Broadcast")


class _Unary(_Elementwise):
    """f(a); ``d(a)`` returns f'(a) as a formula (``VectorizedScalarOp.DiffT``,
    ``formulas/VectorizedScalarOp.py:39-47``)."""

    def difft(self, v, g):
        (a,) = self.children
        d = self.d(a)
        if d is None:
            return Zero(v.dim)
        return a.difft(v, g * d)


class _Square(_Unary):
    def f(self, env, a): return a * a
    def d(self, a): return 2 * a                                                 # maths/Square.py:20-22



def Square(a):
    return a if a.iszero() else _Square(a)

class Exp(_Unary):
    def f(self, env, a): return np.exp(a)
    def d(self, a): return self                                                  # maths/Exp.py:17-19


class Log(_Unary):
    def f(self, env, a): return np.log(a)
    def d(self, a): return Inv(a)


class Inv(_Unary):
    def f(self, env, a): return env.dtype.type(1) / a
    def d(self, a): return -Inv(Square(a))                                       # maths/Inv.py


class Sqrt(_Unary):
    def f(self, env, a): return np.sqrt(a)
    def d(self, a): return IntInv(2) * Rsqrt(a)                                  # maths/Sqrt.py


class Rsqrt(_Unary):
    def f(self, env, a):                                                         # utils/math_functions.py:104-108
        with np.errstate(divide="ignore"):
            return np.where(a == 0, env.dtype.type(0), env.dtype.type(1) / np.sqrt(a))
    def d(self, a): return IntInv(-2) * Pow(self, 3)                             # maths/Rsqrt.py


class Sin(_Unary):
    def f(self, env, a): return np.sin(a)
    def d(self, a): return Cos(a)


class Cos(_Unary):
    def f(self, env, a): return np.cos(a)
    def d(self, a): return -Sin(a)


class Abs(_Unary):
    def f(self, env, a): return np.abs(a)
    def d(self, a): return Sign(a)


class Sign(_Unary):
    def f(self, env, a): return np.sign(a)
    def d(self, a): return None


class Step(_Unary):
    def f(self, env, a): return np.where(a < 0, env.dtype.type(0), env.dtype.type(1))
    def d(self, a): return None


class ReLU(_Unary):
    def f(self, env, a): return np.where(a < 0, env.dtype.type(0), a)
    def d(self, a): return Step(a)


class Floor(_Unary):
    def f(self, env, a): return np.floor(a)
    def d(self, a): return None


class Atan(_Unary):
    def f(self, env, a): return np.arctan(a)
    def d(self, a): return Inv(1 + Square(a))


class Asin(_Unary):
    def f(self, env, a): return np.arcsin(a)
    def d(self, a): return Rsqrt(1 - Square(a))


class Acos(_Unary):
    def f(self, env, a): return np.arccos(a)
    def d(self, a): return -Rsqrt(1 - Square(a))


class XLogX(_Unary):
    def f(self, env, a):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(a != 0, a * np.log(a), env.dtype.type(0))
    def d(self, a): return Log(a) + 1


class Pow(_Unary):
    def __init__(self, a, m):
        super().__init__(a)
        self.m = int(m)
    def f(self, env, a): return np.power(a, self.m).astype(env.dtype)
    def d(self, a): return self.m * Pow(a, self.m - 1) if self.m != 1 else Cst(1)


class Powf(_Elementwise):
    # reference quirk: ``keops_powf`` is the C *float* function ``powf`` for every dtype
    # (``utils/math_functions.py:97-99``), so fp64 formulas get a float32-accurate power.
    def f(self, env, a, b): return np.power(a.astype(np.float32), b.astype(np.float32)).astype(env.dtype)
    def difft(self, v, g):
        a, b = self.children                                                     # maths/Powf.py
        return a.difft(v, _red(g * (b * Powf(a, b - 1)), a.dim)) + b.difft(v, _red(g * (Log(a) * self), b.dim))


class LessThan(_Elementwise):
    def f(self, env, a, b): return (a < b).astype(env.dtype)
    def difft(self, v, g): return Zero(v.dim)


class LessOrEqual(_Elementwise):
    def f(self, env, a, b): return (a <= b).astype(env.dtype)
    def difft(self, v, g): return Zero(v.dim)


class Equal(_Elementwise):
    def f(self, env, a, b): return (a == b).astype(env.dtype)                    # maths/Equal.py
    def difft(self, v, g): return Zero(v.dim)


class NotEqual(_Elementwise):
    def f(self, env, a, b): return (a != b).astype(env.dtype)                    # maths/NotEqual.py
    def difft(self, v, g): return Zero(v.dim)


class Atan2(_Elementwise):
    def f(self, env, a, b): return np.arctan2(a, b)                              # math_functions.py:85
    def difft(self, v, g):                                                       # maths/Atan2.py: d atan2(y, x) = (x dy - y dx) / (x^2 + y^2)
        a, b = self.children
        r2 = Square(a) + Square(b)
        return a.difft(v, _red(g * b / r2, a.dim)) - b.difft(v, _red(g * a / r2, b.dim))


class Mod(_Elementwise):
    def f(self, env, x, n, d): return x - n * np.floor((x - d) / n)              # math_functions.py:157-160
    def difft(self, v, g):                                                       # maths/Mod.py:14-20: gradient of the
        x, n, d = self.children                                                  # alternative form x - n*Floor((x-d)/n)
        return (x - n * Floor((x - d) / n)).difft(v, g)


class SinXDivX(_Unary):
    def f(self, env, a):                                                         # math_functions.py:78-81
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(a != 0, np.sin(a) / np.where(a != 0, a, env.dtype.type(1)), env.dtype.type(1))
    def d(self, a): return Cos(a) / a - Sin(a) / Square(a)                       # maths/SinXDivX.py


class _VecReduce(Op):
    """Max / Min / ArgMax / ArgMin of a vector: sequential scan with a strict comparison, so the first extremum wins
    (maths/Max.py:25-38, maths/ArgMax.py:29-47)."""
    arg, sign = False, 1

    def __init__(self, a):
        self.children = (a,)
        self.dim = 1

    def ev(self, env):
        a = self.sign * self.children[0].ev(env)
        best, idx = a[..., 0:1].copy(), np.zeros_like(a[..., 0:1])
        for k in range(1, a.shape[-1]):
            upd = a[..., k:k + 1] > best
            best = np.where(upd, a[..., k:k + 1], best)
            idx = np.where(upd, env.dtype.type(k), idx)
        return idx if self.arg else self.sign * best

    def difft(self, v, g):
        if self.arg:
            return Zero(v.dim)
        a = self.children[0]                                                     # maths/Max.py: gradient through the arg-extremum
        return a.difft(v, OneHot((ArgMax if self.sign == 1 else ArgMin)(a), a.dim) * g)


class Max(_VecReduce):
    pass


class Min(_VecReduce):
    sign = -1


class ArgMax(_VecReduce):
    arg = True


class ArgMin(_VecReduce):
    arg, sign = True, -1


class OneHot(Op):
    """``out[(int)(a + .5)] = 1`` (maths/OneHot.py:25-37)"""

    def __init__(self, a, dim):
        self.children, self.dim = (a,), int(dim)

    def ev(self, env):
        a = self.children[0].ev(env)
        k = (a + 0.5).astype(np.int64)
        return (k == np.arange(self.dim).reshape((1,) * (a.ndim - 1) + (self.dim,))).astype(env.dtype)

    def difft(self, v, g):
        return Zero(v.dim)


class IfElse(_Elementwise):
    def f(self, env, c, a, b): return np.where(c >= 0, a, b)                     # math_functions.py keops_ifelse
    def difft(self, v, g):
        c, a, b = self.children
        z = Zero(g.dim)
        return a.difft(v, _red(IfElse(c, g, z), a.dim)) + b.difft(v, _red(IfElse(c, z, g), b.dim))


class _Sum(Op):

    def __init__(self, a):
        self.children = (a,)
        self.dim = 1

    def ev(self, env):
        a = self.children[0].ev(env)
        acc = a[..., 0:1].copy()                 # out = 0; out += a[k], k = 0..dim-1
        for k in range(1, a.shape[-1]):
            acc = acc + a[..., k:k + 1]
        return acc

    def difft(self, v, g):
        a = self.children[0]
        return a.difft(v, SumT(g, a.dim))                                        # maths/Sum.py:26-30



def Sum(a):
    if a.dim == 1:
        return a
    if a.iszero():
        return Zero(1)
    return _Sum(a)

class _SumT(Op):

    def __init__(self, a, dim):
        self.children = (a,)
        self.dim = int(dim)

    def ev(self, env):
        a = self.children[0].ev(env)
        return np.broadcast_to(a, a.shape[:-1] + (self.dim,))

    def difft(self, v, g):
        return self.children[0].difft(v, Sum(g))                                 # maths/SumT.py:28-32



def SumT(a, dim):
    if dim == 1:
        return a
    if a.iszero():
        return Zero(dim)
    return _SumT(a, dim)

class _Scalprod(Op):

    def __init__(self, a, b):
        if a.dim != b.dim:
            raise ValueError("Scalprod: dimensions must be the same")
        self.children = (a, b)
        self.dim = 1

    def ev(self, env):
        a, b = [c.ev(env) for c in self.children]
        acc = a[..., 0:1] * b[..., 0:1]
        for k in range(1, self.children[0].dim):
            acc = acc + a[..., k:k + 1] * b[..., k:k + 1]
        return acc

    def difft(self, v, g):
        a, b = self.children                                                     # maths/Scalprod.py:38-40
        return a.difft(v, g * b) + b.difft(v, g * a)



def Scalprod(a, b):
    if _z(a, b):
        return Zero(1)
    return _Scalprod(a, b)

class Concat(Op):
    def __init__(self, a, b):
        self.children = (a, b)
        self.dim = a.dim + b.dim

    def ev(self, env):
        a, b = [c.ev(env) for c in self.children]
        shp = np.broadcast_shapes(a.shape[:-1], b.shape[:-1])
        return np.concatenate([np.broadcast_to(a, shp + a.shape[-1:]), np.broadcast_to(b, shp + b.shape[-1:])], -1)

    def difft(self, v, g):
        a, b = self.children                                                     # maths/Concat.py:28-33
        return a.difft(v, Extract(g, 0, a.dim)) + b.difft(v, Extract(g, a.dim, b.dim))


class _Extract(Op):

    def __init__(self, a, start, dim):
        self.children = (a,)
        self.start, self.dim = int(start), int(dim)

    def ev(self, env):
        return self.children[0].ev(env)[..., self.start:self.start + self.dim]

    def difft(self, v, g):
        a = self.children[0]
        return a.difft(v, ExtractT(g, self.start, a.dim))                        # maths/Extract.py:31-35



def Extract(a, start, dim):
    if a.iszero():
        return Zero(dim)
    return _Extract(a, start, dim)

class _ExtractT(Op):

    def __init__(self, a, start, dim):
        self.children = (a,)
        self.start, self.dim = int(start), int(dim)

    def ev(self, env):
        a = self.children[0].ev(env)
        out = np.zeros(a.shape[:-1] + (self.dim,), env.dtype)
        out[..., self.start:self.start + a.shape[-1]] = a
        return out

    def difft(self, v, g):
        a = self.children[0]
        return a.difft(v, Extract(g, self.start, a.dim))



def ExtractT(a, start, dim):
    if a.iszero():
        return Zero(dim)
    return _ExtractT(a, start, dim)

class Round(_Unary):
    def __init__(self, a, d=0):
        super().__init__(a)
        self.d = int(d)
    def f(self, env, a):                                                         # math_functions.py keops_round
        r = lambda t: np.where(t >= 0, np.floor(t + 0.5), np.ceil(t - 0.5))     # C round(): half away from zero
        if self.d == 0:
            return r(a).astype(env.dtype)
        s = env.dtype.type(10 ** self.d)
        return (r(a * s) / s).astype(env.dtype)
    def d_(self, a): return None
    def difft(self, v, g): return Zero(v.dim)


class Clamp(_Elementwise):
    def f(self, env, x, a, b): return np.where(x < a, a, np.where(x > b, b, x))  # math_functions.py keops_clamp
    def difft(self, v, g):
        x, a, b = self.children                                                  # maths/Clamp.py (ClampInt form)
        alt = ClampInt((x - a) / (b - a), 0, 1) * (b - a) + a
        return alt.difft(v, g)


class ClampInt(_Unary):
    def __init__(self, a, lo, hi):
        super().__init__(a)
        self.lo, self.hi = int(lo), int(hi)
    def f(self, env, a): return np.clip(a, env.dtype.type(self.lo), env.dtype.type(self.hi))
    def d(self, a): return DiffClampInt(a, self.lo, self.hi)


class DiffClampInt(_Unary):
    def __init__(self, a, lo, hi):
        super().__init__(a)
        self.lo, self.hi = int(lo), int(hi)
    def f(self, env, a): return np.where(a < self.lo, 0, np.where(a > self.hi, 0, 1)).astype(env.dtype)
    def d(self, a): return None


class TensorProd(Op):
    def __init__(self, a, b):
        self.children = (a, b)
        self.dim = a.dim * b.dim
    def ev(self, env):
        a, b = [c.ev(env) for c in self.children]
        return (a[..., :, None] * b[..., None, :]).reshape(np.broadcast_shapes(a.shape[:-1], b.shape[:-1]) + (self.dim,))
    def difft(self, v, g):
        a, b = self.children                                                     # maths/TensorProd.py
        return a.difft(v, MatVecMult(g, b)) + b.difft(v, VecMatMult(a, g))


class MatVecMult(Op):
    def __init__(self, A, b):
        self.children = (A, b)
        self.dim = A.dim // b.dim
    def ev(self, env):
        A, b = [c.ev(env) for c in self.children]
        m = self.children[1].dim
        A = A.reshape(A.shape[:-1] + (self.dim, m))
        acc = A[..., 0] * b[..., 0:1]
        for k in range(1, m):
            acc = acc + A[..., k] * b[..., k:k + 1]
        return acc
    def difft(self, v, g):
        A, b = self.children                                                     # maths/MatVecMult.py
        return A.difft(v, TensorProd(g, b)) + b.difft(v, VecMatMult(g, A))


class VecMatMult(Op):
    def __init__(self, b, A):
        self.children = (b, A)
        self.dim = A.dim // b.dim
    def ev(self, env):
        b, A = [c.ev(env) for c in self.children]
        m = self.children[0].dim
        A = A.reshape(A.shape[:-1] + (m, self.dim))
        acc = b[..., 0:1] * A[..., 0, :]
        for k in range(1, m):
            acc = acc + b[..., k:k + 1] * A[..., k, :]
        return acc
    def difft(self, v, g):
        b, A = self.children
        return A.difft(v, TensorProd(b, g)) + b.difft(v, MatVecMult(A, g))


class TensorDot(Op):
    """``maths/TensorDot.py``: contraction of two flat row-major tensors over paired axes.  The reference's loop nest
    (``:155-187``) runs over the kept axes of ``a``, then of ``b`` (outer), then the contracted axes (inner) and does
    ``out[o] += a[ia] * b[ib]``; kept axis number ``t`` (in that order) lands on axis ``permute[t]`` of the output."""
    def __init__(self, a, b, dimsfa, dimsfb, contfa, contfb, permute=None):
        self.children = (a, b)
        self.da, self.db = [int(d) for d in dimsfa], [int(d) for d in dimsfb]
        self.ca, self.cb = [int(c) for c in contfa], [int(c) for c in contfb]
        assert a.dim == int(np.prod(self.da, dtype=np.int64)) and b.dim == int(np.prod(self.db, dtype=np.int64))
        assert [self.da[c] for c in self.ca] == [self.db[c] for c in self.cb]
        self.ka = [k for k in range(len(self.da)) if k not in self.ca]
        self.kb = [k for k in range(len(self.db)) if k not in self.cb]
        kept = [self.da[k] for k in self.ka] + [self.db[k] for k in self.kb]
        self.perm = list(range(len(kept))) if permute is None else [int(q) for q in permute]
        assert sorted(self.perm) == list(range(len(kept)))
        self.oshape = [kept[self.perm.index(q)] for q in range(len(kept))]
        self.dim = int(np.prod(kept, dtype=np.int64))
        # flat index tables over the loop nest (kept..., contracted...)
        loop = kept + [self.da[c] for c in self.ca]
        grid = np.indices(loop).reshape(len(loop), -1) if loop else np.zeros((0, 1), dtype=np.int64)
        nka, nk = len(self.ka), len(kept)
        ia = [None] * len(self.da)
        for t, ax in enumerate(self.ka): ia[ax] = grid[t]
        for c, ax in enumerate(self.ca): ia[ax] = grid[nk + c]
        ib = [None] * len(self.db)
        for t, ax in enumerate(self.kb): ib[ax] = grid[nka + t]
        for c, ax in enumerate(self.cb): ib[ax] = grid[nk + c]
        io = [grid[self.perm.index(q)] for q in range(nk)]
        flat = lambda idx, shape: np.ravel_multi_index(idx, shape) if shape else np.zeros(grid.shape[1], dtype=np.int64)
        self.ia, self.ib, self.io = flat(ia, self.da), flat(ib, self.db), flat(io, self.oshape)
    def ev(self, env):
        a, b = [c.ev(env) for c in self.children]
        lead = np.broadcast_shapes(a.shape[:-1], b.shape[:-1])
        out = np.zeros(lead + (self.dim,), dtype=env.dtype)
        for o, i, j in zip(self.io, self.ia, self.ib):                           # the reference's order of additions
            out[..., o] += a[..., i] * b[..., j]
        return out
    def difft(self, v, g):
        a, b = self.children                                                     # TensorDot.py:105-139,189-216
        nka = len(self.ka)
        order = lambda keys: sorted(range(len(keys)), key=keys.__getitem__)
        back_a = [self.ka[t] for t in order(self.perm[:nka])] + [self.ca[c] for c in order(self.cb)]
        back_b = [self.kb[t] for t in order(self.perm[nka:])] + [self.cb[c] for c in order(self.ca)]
        return a.difft(v, TensorDot(g, b, self.oshape, self.db, self.perm[nka:], self.kb, back_a)) + \
            b.difft(v, TensorDot(g, a, self.oshape, self.da, self.perm[:nka], self.ka, back_b))


def Kron(a, b, dimsfa, dimsfb):                                                  # maths/Kron.py:17-31
    n = len(dimsfa)
    return TensorDot(a, b, dimsfa, dimsfb, [], [], list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2)))


def _softdtw_forward(env, x, y, gam):
    """All prefix values r[j][i] of ``maths/SoftDTW_SqDist.py:33-93`` (same recursion as ``:132-157``), on (M, N) arrays."""
    n, m = x.shape[-1], y.shape[-1]
    X = lambda i: x[..., i]
    Y = lambda j: y[..., j]
    r = [[None] * n for _ in range(m)]
    d = X(0) - Y(0)
    r[0][0] = d * d
    for i in range(1, n):
        d = X(i) - Y(0)
        r[0][i] = d * d + r[0][i - 1]
    for j in range(1, m):
        d = X(0) - Y(j)
        r[j][0] = d * d + r[j - 1][0]
        for i in range(1, n):
            d = X(i) - Y(j)
            a, b, c = r[j - 1][i - 1], r[j - 1][i], r[j][i - 1]
            # MIN3 goes through fminf, i.e. float32 even for double data; it only shifts the log-sum-exp
            mn = np.minimum(np.minimum(a.astype(np.float32), b.astype(np.float32)), c.astype(np.float32)).astype(env.dtype)
            r[j][i] = d * d + (mn - gam * np.log(np.exp((mn - a) / gam) + np.exp((mn - c) / gam) + np.exp((mn - b) / gam)))
    return r


class SoftDTW_SqDist(Op):
    def __init__(self, x, y, gamma):
        assert gamma.dim == 1
        self.children = (x, y, gamma)
        self.dim = 1
    def ev(self, env):
        x, y, gam = [c.ev(env) for c in self.children]
        r = _softdtw_forward(env, x, y, gam[..., 0])
        return r[-1][-1][..., None]
    def difft(self, v, g):
        x, y, gamma = self.children                                               # SoftDTW_SqDist.py:95-106
        assert v.ind not in gamma.vars(), "autograd wrt gamma in SoftDTW_SqDist operation not implemented."
        full = GradSoftDTW_SqDist(x, y, gamma) * g
        return x.difft(v, Extract(full, 0, x.dim)) + y.difft(v, Extract(full, x.dim, y.dim))


class GradSoftDTW_SqDist(Op):
    def __init__(self, x, y, gamma):
        self.children = (x, y, gamma)
        self.dim = x.dim + y.dim
    def ev(self, env):                                                            # SoftDTW_SqDist.py:124-211
        x, y, gam = [c.ev(env) for c in self.children]
        gam = gam[..., 0]
        n, m = x.shape[-1], y.shape[-1]
        r = _softdtw_forward(env, x, y, gam)
        X = lambda i: x[..., i]
        Y = lambda j: y[..., j]
        gx, gy = [None] * n, [None] * m
        ejp1 = [None] * n
        eij = np.ones(np.broadcast_shapes(x.shape[:-1], y.shape[:-1]), dtype=env.dtype)
        eip1j = eij
        d = X(n - 1) - Y(m - 1)
        gx[n - 1] = 2 * d * eij
        gy[m - 1] = -2 * d * eij
        for i in range(n - 2, -1, -1):
            a = np.exp((r[m - 1][i + 1] - r[m - 1][i] - d * d) / gam)
            eij = a * eip1j
            ejp1[i + 1] = eip1j
            eip1j = eij
            d = X(i) - Y(m - 1)
            gx[i] = 2 * d * eij
            gy[m - 1] = gy[m - 1] - 2 * d * eij
        ejp1[0] = eij
        for j in range(m - 2, -1, -1):
            d = X(n - 1) - Y(j + 1)
            b = np.exp((r[j + 1][n - 1] - r[j][n - 1] - d * d) / gam)
            eij = b * ejp1[n - 1]
            eip1j = eij
            d = X(n - 1) - Y(j)
            gx[n - 1] = gx[n - 1] + 2 * d * eij
            gy[j] = -2 * d * eij
            for i in range(n - 2, -1, -1):
                a = np.exp((r[j][i + 1] - r[j][i] - d * d) / gam)
                d = X(i) - Y(j + 1)
                b = np.exp((r[j + 1][i] - r[j][i] - d * d) / gam)
                d = X(i + 1) - Y(j + 1)
                c = np.exp((r[j + 1][i + 1] - r[j][i] - d * d) / gam)
                eij = a * eip1j + b * ejp1[i] + c * ejp1[i + 1]
                ejp1[i + 1] = eip1j
                eip1j = eij
                d = X(i) - Y(j)
                gx[i] = gx[i] + 2 * d * eij
                gy[j] = gy[j] - 2 * d * eij
            ejp1[0] = eij
        shape = np.broadcast_shapes(x.shape[:-1], y.shape[:-1])
        return np.stack([np.broadcast_to(t, shape) for t in gx + gy], axis=-1).astype(env.dtype)
    def difft(self, v, g):
        raise ValueError("autograd for GradSoftDTW_SqDist operation not implemented.")


class BSpline(Op):
    """``maths/BSpline.py:19-126``: one-hot of the knot span containing x, then ``order`` sweeps of
    ``out[i] = w(i,k) out[i] + (1 - w(i+1,k)) out[i+1]`` with ``w(i,k) = (x - t[i]) / (t[i+k] - t[i])`` (0 if the span is empty);
    the reference keeps x and the ratios in C ``float`` -- and does not generate at all for double inputs."""
    def __init__(self, knots, x, order):
        order = int(order)
        assert order >= 0 and x.dim == 1 and knots.dim >= order + 2
        self.children = (knots, x)
        self.order = order
        self.dim = knots.dim - order - 1
    def ev(self, env):
        t, x = [c.ev(env) for c in self.children]
        x = x[..., 0]
        T = lambda i: t[..., i]
        nb = t.shape[-1] - 1
        out = [((T(j) <= x) & (x < T(j + 1))).astype(env.dtype) for j in range(nb)]
        def w(i, k):
            with np.errstate(divide="ignore", invalid="ignore"):
                return np.where(T(i) < T(i + k), (x - T(i)) / (T(i + k) - T(i)), env.dtype.type(0)).astype(env.dtype)
        for k in range(1, self.order + 1):
            r1 = w(0, k)
            for i in range(nb - k):
                r2 = w(i + 1, k)
                out[i] = r1 * out[i] + (env.dtype.type(1) - r2) * out[i + 1]
                r1 = r2
        shape = np.broadcast_shapes(*[o.shape for o in out[:self.dim]])
        return np.stack([np.broadcast_to(o, shape) for o in out[:self.dim]], axis=-1)
    def difft(self, v, g):
        raise NotImplementedError("KeOps BSplines are not yet differentiable.")


# ---- complex arithmetic on interleaved (re, im) pairs: keopscore/keopscore/formulas/complex/*.py ----------------------------
def _cdims(*ops):
    dims = set(o.dim for o in ops)
    assert max(dims) % 2 == 0 and len(dims) <= 2 and (len(dims) == 1 or min(dims) == 2), "VectorizedComplexScalarOp dims"
    return max(dims)


def _cpairs(a, dim):
    """(..., 2) or (..., dim) interleaved -> (re, im) arrays of trailing size dim // 2 (a single complex number broadcasts)"""
    return a[..., 0::2], a[..., 1::2]


def _cjoin(re, im):
    shape = np.broadcast_shapes(re.shape, im.shape)
    out = np.empty(shape[:-1] + (2 * shape[-1],), dtype=np.result_type(re, im))
    out[..., 0::2], out[..., 1::2] = re, im
    return out


class ComplexReal(Op):
    def __init__(self, f):
        assert f.dim % 2 == 0
        self.children, self.dim = (f,), f.dim // 2
    def ev(self, env): return self.children[0].ev(env)[..., 0::2]
    def difft(self, v, g): return self.children[0].difft(v, Real2Complex(g))


class ComplexImag(Op):
    def __init__(self, f):
        assert f.dim % 2 == 0
        self.children, self.dim = (f,), f.dim // 2
    def ev(self, env): return self.children[0].ev(env)[..., 1::2]
    def difft(self, v, g): return self.children[0].difft(v, Imag2Complex(g))


class Real2Complex(Op):
    def __init__(self, f): self.children, self.dim = (f,), 2 * f.dim
    def ev(self, env):
        a = self.children[0].ev(env)
        return _cjoin(a, np.zeros_like(a))
    def difft(self, v, g): return self.children[0].difft(v, ComplexReal(g))


class Imag2Complex(Op):
    def __init__(self, f): self.children, self.dim = (f,), 2 * f.dim
    def ev(self, env):
        a = self.children[0].ev(env)
        return _cjoin(np.zeros_like(a), a)
    def difft(self, v, g): return self.children[0].difft(v, ComplexImag(g))


class Conj(Op):
    def __init__(self, f): self.children, self.dim = (f,), _cdims(f)
    def ev(self, env):
        re, im = _cpairs(self.children[0].ev(env), self.dim)
        return _cjoin(re, -im)
    def difft(self, v, g): return self.children[0].difft(v, Conj(g))


class ComplexSquareAbs(Op):
    def __init__(self, f):
        assert f.dim % 2 == 0
        self.children, self.dim = (f,), f.dim // 2
    def ev(self, env):
        re, im = _cpairs(self.children[0].ev(env), 2 * self.dim)
        return re * re + im * im
    def difft(self, v, g):
        f = self.children[0]                                                      # ComplexSquareAbs.py: alternative formula
        return ComplexReal(ComplexMult(f, Conj(f))).difft(v, g)


def ComplexAbs(f): return Sqrt(ComplexSquareAbs(f))
def ComplexAngle(f): return Atan2(ComplexImag(f), ComplexReal(f))


class _ComplexBinary(Op):
    def __init__(self, f, g):
        self.children, self.dim = (f, g), _cdims(f, g)
    def ev(self, env):
        a, b = [c.ev(env) for c in self.children]
        return _cjoin(*self.f(_cpairs(a, self.dim), _cpairs(b, self.dim)))
    def _split(self, g):
        f, h = self.children
        gf = ComplexSum(g) if (f.dim == 2 and h.dim > 2) else g
        gh = ComplexSum(g) if (h.dim == 2 and f.dim > 2) else g
        return gf, gh


class ComplexAdd(_ComplexBinary):
    def f(self, a, b): return a[0] + b[0], a[1] + b[1]
    def difft(self, v, g):
        gf, gh = self._split(g)
        return self.children[0].difft(v, gf) + self.children[1].difft(v, gh)


class ComplexSubtract(_ComplexBinary):
    def f(self, a, b): return a[0] - b[0], a[1] - b[1]
    def difft(self, v, g):
        gf, gh = self._split(g)
        return self.children[0].difft(v, gf) - self.children[1].difft(v, gh)


class ComplexMult(_ComplexBinary):
    def f(self, a, b): return a[0] * b[0] - a[1] * b[1], a[0] * b[1] + a[1] * b[0]
    def difft(self, v, g):
        f, h = self.children                                                      # ComplexMult.py:30-50
        gf, gh = ComplexMult(Conj(h), g), ComplexMult(Conj(f), g)
        if f.dim == 2 and h.dim > 2:
            gf = ComplexSum(gf)
        elif h.dim == 2 and f.dim > 2:
            gh = ComplexSum(gh)
        return f.difft(v, gf) + h.difft(v, gh)


def ComplexDivide(f, g): return ComplexMult(Real2Complex(Inv(ComplexSquareAbs(g))), ComplexMult(f, Conj(g)))


class _ComplexRealScal(Op):
    def __init__(self, f, g):
        assert f.dim == 1 and g.dim % 2 == 0
        self.children, self.dim = (f, g), g.dim
    def ev(self, env):
        a, b = [c.ev(env) for c in self.children]
        return a * b
    def difft(self, v, g):
        f, h = self.children
        return ComplexMult(Real2Complex(f), h).difft(v, g)


def ComplexRealScal(f, g): return _ComplexRealScal(f, g) if f.dim == 1 else _ComplexRealScal(g, f)


class ComplexExp(Op):
    def __init__(self, f): self.children, self.dim = (f,), _cdims(f)
    def ev(self, env):
        re, im = _cpairs(self.children[0].ev(env), self.dim)
        r = np.exp(re)
        return _cjoin(r * np.cos(im), r * np.sin(im))
    def difft(self, v, g):
        f = self.children[0]                                                      # ComplexExp.py: alternative formula
        mod = Exp(ComplexReal(f))
        alt = Real2Complex(mod * Cos(ComplexImag(f))) + Imag2Complex(mod * Sin(ComplexImag(f)))
        return alt.difft(v, g)


class ComplexExp1j(Op):
    def __init__(self, f): self.children, self.dim = (f,), 2 * f.dim
    def ev(self, env):
        a = self.children[0].ev(env)
        return _cjoin(np.cos(a), np.sin(a))
    def difft(self, v, g):
        f = self.children[0]
        return (Real2Complex(Cos(f)) + Imag2Complex(Sin(f))).difft(v, g)


class ComplexSum(Op):
    def __init__(self, f):
        assert f.dim % 2 == 0
        self.children, self.dim = (f,), 2
    def ev(self, env):
        a = self.children[0].ev(env)
        re, im = np.zeros(a.shape[:-1], env.dtype), np.zeros(a.shape[:-1], env.dtype)
        for k in range(0, a.shape[-1], 2):
            re, im = re + a[..., k], im + a[..., k + 1]
        return np.stack([re, im], axis=-1)
    def difft(self, v, g): return self.children[0].difft(v, ComplexSumT(g, self.children[0].dim))


class ComplexSumT(Op):
    def __init__(self, f, dim):
        assert f.dim == 2
        self.children, self.dim = (f,), int(dim)
    def ev(self, env): return np.tile(self.children[0].ev(env), self.dim // 2)
    def difft(self, v, g): return self.children[0].difft(v, ComplexSum(g))


# ---- symbolic operators on a formula that is linear in a placeholder variable (formulas/LinearOperators/*.py).  Restated
# numerically: the matrix M of v -> f(v) is read off column by column by evaluating f at the basis vectors e_k.
class _EnvWith:
    def __init__(self, env, ind, arr):
        self._env, self._ind, self._arr = env, ind, arr
    def __getattr__(self, name):
        return getattr(self._env, name)
    def var(self, v):
        return self._arr if v.ind == self._ind else self._env.var(v)


class _LinearOperator(Op):
    def _columns(self, env):
        f, v = self.f, self.v
        for k in range(v.dim):
            e_k = np.zeros((1, 1, v.dim), dtype=env.dtype)
            e_k[..., k] = 1
            yield k, f.ev(_EnvWith(env, v.ind, e_k))                              # column k of M, shape (..., f.dim)
    def vars(self):
        out = dict(self.f.vars())
        out.pop(self.v.ind, None)
        for c in self.children:
            out.update(c.vars())
        return out
    def difft(self, v, g):
        raise ValueError("differentiation through TraceOperator / AdjointOperator is not restated in the oracle")


class _TraceOperator(_LinearOperator):
    """TraceOperator.py:30-52: f.dim == v.dim: tr(diag(w) M); f.dim == 1: sum_i w_i M_i (w scalar or of size v.dim)."""
    def __init__(self, f, v, w=1):
        self.f, self.v = f, v
        self.w = None if (isinstance(w, int) and w == 1) else _w(w)
        assert f.dim in (1, v.dim) and (self.w is None or self.w.dim in (1, v.dim))
        self.children = () if self.w is None else (self.w,)
        self.dim = 1
    def ev(self, env):
        w = None if self.w is None else self.w.ev(env)
        total = 0
        for k, col in self._columns(env):
            term = col[..., k] if self.f.dim == self.v.dim and self.v.dim > 1 else col[..., 0]
            if w is not None:
                term = term * (w[..., k] if w.shape[-1] > 1 else w[..., 0])
            total = total + term
        return np.asarray(total, dtype=env.dtype)[..., None]


class _AdjointOperator(_LinearOperator):
    """AdjointOperator.py: (M^T u)_k = <column k of M, u>."""
    def __init__(self, f, v, u):
        self.f, self.v, self.u = f, v, _w(u)
        assert self.u.dim == f.dim
        self.children = (self.u,)
        self.dim = v.dim
    def ev(self, env):
        u = self.u.ev(env)
        cols = [(col * u).sum(-1) for _, col in self._columns(env)]
        shape = np.broadcast_shapes(*[c.shape for c in cols])
        return np.stack([np.broadcast_to(c, shape) for c in cols], axis=-1).astype(env.dtype)


def TraceOperator(f, v, w=1): return _TraceOperator(f, v, w)
def AdjointOperator(f, v, u): return _AdjointOperator(f, v, u)


def Elem(a, m): return Extract(a, m, 1)
def ElemT(a, n, m): return ExtractT(a, m, n)
def SqNorm2(a): return Sum(Square(a))                                            # maths/SqNorm2.py
def Norm2(a): return Sqrt(Scalprod(a, a))                                        # maths/Norm2.py
def Normalize(a): return Rsqrt(SqNorm2(a)) * a                                   # maths/Normalize.py
def SqDist(a, b): return SqNorm2(a - b)                                          # maths/SqDist.py


def WeightedSqNorm(A, x):
    if A.dim == 1:
        return A * SqNorm2(x)
    if A.dim == x.dim:
        return Sum(A * Square(x))
    raise NotImplementedError("SymSqNorm is not restated in the oracle")


def WeightedSqDist(A, x, y): return WeightedSqNorm(A, x - y)


# ------------------------------------------------------------------------------------------
# reductions
# ------------------------------------------------------------------------------------------
class Reduction:
    def __init__(self, kind, F, tagI, K=1, G=None):
        self.kind, self.F, self.tagI, self.K, self.G = kind, F, int(tagI), int(K), G

    def vars(self):
        out = dict(self.F.vars())
        if self.G is not None:
            out.update(self.G.vars())
        return out


def Sum_Reduction(F, tagI): return Reduction("Sum", F, tagI)
def Zero_Reduction(dim, tagI): return Reduction("Zero", Zero(dim), tagI)
def Max_SumShiftExpWeight_Reduction(F, tagI, G=None): return Reduction("LSE", F, tagI, G=_w(1 if G is None else G))
Max_SumShiftExp_Reduction = Max_SumShiftExpWeight_Reduction
def Min_Reduction(F, tagI): return Reduction("Min", F, tagI)
def Max_Reduction(F, tagI): return Reduction("Max", F, tagI)
def ArgMin_Reduction(F, tagI): return Reduction("ArgMin", F, tagI)
def ArgMax_Reduction(F, tagI): return Reduction("ArgMax", F, tagI)
def Min_ArgMin_Reduction(F, tagI): return Reduction("Min_ArgMin", F, tagI)
def Max_ArgMax_Reduction(F, tagI): return Reduction("Max_ArgMax", F, tagI)
def KMin_Reduction(F, K, tagI): return Reduction("KMin", F, tagI, K)
def ArgKMin_Reduction(F, K, tagI): return Reduction("ArgKMin", F, tagI, K)
def KMin_ArgKMin_Reduction(F, K, tagI): return Reduction("KMin_ArgKMin", F, tagI, K)


def Grad_WithSavedForward(red, v, gradin, f0):
    """``reductions/Sum_Reduction.py:42-45``, ``Max_SumShiftExpWeight_Reduction.py:89-103``,
    ``ArgKMin_Reduction.py:33-37``."""
    newtag = v.cat % 2
    if red.kind == "Sum":
        return Sum_Reduction(red.F.difft(v, gradin), newtag)
    if red.kind == "LSE":
        M = Extract(f0, 0, 1)
        gS = Extract(gradin, 1, red.G.dim)
        H = Exp(red.F - M) * red.G
        return Sum_Reduction(H.difft(v, _red(gS, H.dim) if gS.dim != H.dim else gS), newtag)
    if red.kind in ("ArgMin", "ArgMax", "ArgKMin", "Zero"):
        return Zero_Reduction(v.dim, newtag)
    raise NotImplementedError("no autodiff through " + red.kind)


def _placeholder(dim, *forms):
    """fresh variable used to take adjoints; it must have vanished from the result (autodiff/Diff.py:24-36)"""
    return Var(10 ** 6 + dim, dim, 2)


def Grad(f, v, gradin):
    """``autodiff/Grad.py:14-27``: formula-level gradient, or the gradient of a reduction."""
    if isinstance(f, Reduction):
        return Grad_WithSavedForward(f, v, gradin, None)
    return f.difft(v, _w(gradin))


def Diff(f, v, u):
    """``autodiff/Diff.py:14-36``: D_v f . u as the adjoint (w.r.t. a placeholder gradin) of the gradient."""
    if isinstance(f, Reduction):
        return Diff_WithSavedForward(f, v, u, None)
    g = _placeholder(f.dim)
    h = f.difft(v, g)
    return (Scalprod(h, u) if v.dim > 1 else h * u).difft(g, IntCst(1))


def Diff_WithSavedForward(red, v, u, f0):
    """``reductions/Sum_Reduction.py:47-50``, ``Max_SumShiftExpWeight_Reduction.py:105-120``."""
    if red.kind == "Sum":
        return Sum_Reduction(Diff(red.F, v, u), red.tagI)
    if red.kind == "LSE":
        M = Extract(f0, 0, 1)
        return Sum_Reduction(Diff(Concat(Zero(1), Exp(red.F - M) * red.G), v, u), red.tagI)
    raise NotImplementedError("no forward-mode autodiff through " + red.kind)


def Divergence(f, v):
    """``autodiff/Divergence.py``: trace of the Jacobian."""
    total = None
    for k in range(v.dim):
        term = Elem(Diff(f, v, ElemT(IntCst(1), v.dim, k)), k)
        total = term if total is None else total + term
    return total


def Laplacian(f, v):
    """``autodiff/Laplacian.py``"""
    return Divergence(f.difft(v, IntCst(1)), v)


def GradMatrix(f, v):
    """``maths/GradMatrix.py:14-28``: gradients of the components of f, concatenated."""
    res = None
    for i in range(f.dim):
        g_i = f.difft(v, ElemT(IntCst(1), f.dim, i))
        res = g_i if res is None else Concat(res, g_i)
    return res


def Factorize(f, g):
    """``formulas/factorization``: evaluation-order hint only, the value is f."""
    return f


def AutoFactorize(f):
    return f


_NAMESPACE = {k: v for k, v in list(globals().items()) if k[0].isupper() and k not in ("Op", "Reduction", "Fraction")}
_NAMESPACE.update({"Vi": lambda p, d: Var(p, d, 0), "Vj": lambda p, d: Var(p, d, 1), "Pm": lambda p, d: Var(p, d, 2)})


def parse(formula, aliases=()):
    local = {}
    for k, alias in enumerate(aliases):
        if alias is None or "=" not in alias:
            continue
        name, rhs = alias.replace(" ", "").split("=")
        m = re.fullmatch(r"(Vi|Vj|Pm)\((\d+)(?:,(\d+))?\)", rhs)
        if m:
            cat = {"Vi": 0, "Vj": 1, "Pm": 2}[m.group(1)]
            pos, dim = (k, int(m.group(2))) if m.group(3) is None else (int(m.group(2)), int(m.group(3)))
            local[name] = Var(pos, dim, cat)
        else:
            local[name] = eval(rhs, dict(_NAMESPACE), {})
    return eval(formula, dict(_NAMESPACE), local)


# ------------------------------------------------------------------------------------------
# evaluation
# ------------------------------------------------------------------------------------------
class _Env:
    def __init__(self, args, dtype, tagI, rows):
        self.args, self.dtype, self.tagI, self.rows = args, np.dtype(dtype), tagI, rows

    def var(self, v):
        a = np.asarray(self.args[v.ind], dtype=self.dtype)
        if v.cat == 2:
            return a.reshape(1, 1, v.dim)
        a = a.reshape(-1, v.dim)
        if v.cat == self.tagI:      # output-side variable: current chunk of rows
            return a[self.rows][:, None, :]
        return a[None, :, :]


def genred(formula, aliases, args, dtype="float32", chunk_elems=1 << 24, sizes=None):
    """Raw kernel output ``(nrow, dim)`` of the reduction string, before any Python post-processing
    (what ``LoadKeOps.genred`` returns, ``pykeops/pykeops/common/keops_io/LoadKeOps.py:177-243``)."""
    red = parse(formula, aliases)
    dtype = np.dtype(dtype)
    vs = red.vars()
    tagI = red.tagI
    nrow = ncol = None
    if sizes is not None:                      # (nx, ny) as handed to the binder: needed when a side has no variable
        nrow, ncol = (sizes[0], sizes[1]) if tagI == 0 else (sizes[1], sizes[0])
    for v in vs.values():
        if v.cat == 2:
            continue
        n = np.asarray(args[v.ind]).reshape(-1, v.dim).shape[0]
        if v.cat == tagI:
            nrow = n
        else:
            ncol = n
    if nrow is None or ncol is None:
        raise ValueError("oracle.genred needs at least one i- and one j-variable (or explicit sizes)")
    F, G, K = red.F, red.G, red.K
    d = F.dim
    kind = red.kind
    dimout = {"Sum": d, "Zero": d, "Min": d, "Max": d, "ArgMin": d, "ArgMax": d, "Min_ArgMin": 2 * d, "Max_ArgMax": 2 * d,
              "KMin": K * d, "ArgKMin": K * d, "KMin_ArgKMin": 2 * K * d}.get(kind)
    if kind == "LSE":
        dimout = 1 + G.dim
    out = np.zeros((nrow, dimout), dtype)
    if kind == "Zero" or nrow == 0:
        return out
    step = max(1, chunk_elems // max(1, ncol * max(d, 1)))
    for r0 in range(0, nrow, step):
        rows = slice(r0, min(nrow, r0 + step))
        nr = rows.stop - rows.start
        env = _Env(args, dtype, tagI, rows)
        Fv = np.broadcast_to(F.ev(env), (nr, ncol, d))
        if kind == "Sum":
            out[rows] = Fv.sum(axis=1, dtype=np.float64).astype(dtype)
        elif kind == "LSE":
            Gv = np.broadcast_to(G.ev(env), (nr, ncol, G.dim))
            if ncol == 0:
                out[rows, 0] = -np.inf
                continue
            m = Fv.max(axis=1)                                    # (nr, 1)
            e = np.exp(Fv - m[:, None, :])
            s = (e * Gv).sum(axis=1, dtype=np.float64)
            out[rows, 0:1] = m
            out[rows, 1:] = s.astype(dtype)
        elif kind in ("Min", "Max", "ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax"):
            if ncol == 0:
                val = np.full((nr, d), np.inf if "Min" in kind else -np.inf, dtype)
                idx = np.zeros((nr, d), np.int64)
            else:
                idx = Fv.argmin(axis=1) if "Min" in kind else Fv.argmax(axis=1)     # first extremum wins
                val = np.take_along_axis(Fv, idx[:, None, :], axis=1)[:, 0, :]
            if kind in ("Min", "Max"):
                out[rows] = val
            elif kind in ("ArgMin", "ArgMax"):
                out[rows] = idx.astype(dtype)
            else:
                out[rows, :d] = val
                out[rows, d:] = idx.astype(dtype)
        else:
            order = np.argsort(Fv, axis=1, kind="stable")[:, :K, :]                 # ties -> smallest index
            vals = np.take_along_axis(Fv, order, axis=1)
            kk = order.shape[1]
            v_full = np.full((nr, K, d), np.inf, dtype)
            i_full = np.zeros((nr, K, d), dtype)
            v_full[:, :kk] = vals
            i_full[:, :kk] = order.astype(dtype)
            if kind == "KMin":
                out[rows] = v_full.reshape(nr, K * d)
            elif kind == "ArgKMin":
                out[rows] = i_full.reshape(nr, K * d)
            else:
                out[rows] = np.stack([v_full, i_full], axis=2).reshape(nr, 2 * K * d)
    return out


def genred_ranges(formula, aliases, args, ranges, dtype="float32"):
    """Block-sparse reduction (``ranges`` 6-tuple, ``pykeops/pykeops/torch/generic/generic_red.py:581-639``):
    row-range ``k`` is reduced over the union of its column ranges, in increasing column order, exactly
    like the scalar loops of ``CpuReduc_ranges`` (``keopscore/keopscore/mapreduce/cpu/CpuReduc_ranges.py:173-228``);
    arg-type outputs are *global* column indices.  Rows covered by no range are left at 0."""
    red = parse(formula, aliases)
    tagI = red.tagI
    vs = red.vars()
    rx, sx, ry = (np.asarray(r, np.int64) for r in (ranges[0:3] if tagI == 0 else ranges[3:6]))
    nrow = None
    for v in vs.values():
        if v.cat == tagI:
            nrow = np.asarray(args[v.ind]).reshape(-1, v.dim).shape[0]
    out = None
    d, K, kind = red.F.dim, red.K, red.kind
    for k in range(rx.shape[0]):
        r0, r1 = int(rx[k, 0]), int(rx[k, 1])
        if r1 <= r0:
            continue
        s0 = 0 if k == 0 else int(sx[k - 1])
        cols = [np.arange(int(ry[l, 0]), int(ry[l, 1])) for l in range(s0, int(sx[k]))]
        J = np.concatenate(cols) if cols else np.zeros(0, np.int64)
        sub = list(args)
        for v in vs.values():
            a = np.asarray(args[v.ind])
            if v.cat == tagI:
                sub[v.ind] = a.reshape(-1, v.dim)[r0:r1]
            elif v.cat == 1 - tagI:
                sub[v.ind] = a.reshape(-1, v.dim)[J]
        if len(J) == 0:
            continue    # the reference writes the accumulator's initial value here; not exercised by the fixtures
        res = genred(formula, aliases, sub, dtype)
        if out is None:
            out = np.zeros((nrow, res.shape[1]), res.dtype)

        def remap(block):                                   # local column positions -> global column indices
            return J[block.astype(np.int64)].astype(res.dtype)
        if kind in ("ArgMin", "ArgMax"):
            res = remap(res)
        elif kind == "ArgKMin":
            r3 = remap(res).reshape(-1, K, d)
            r3[:, len(J):, :] = 0                           # fewer than K columns: padding entries stay (inf, 0)
            res = r3.reshape(res.shape)
        elif kind in ("Min_ArgMin", "Max_ArgMax"):
            res[:, d:] = remap(res[:, d:])
        elif kind == "KMin_ArgKMin":
            r3 = res.reshape(-1, K, 2, d)
            r3[:, :, 1, :] = np.where(np.isfinite(r3[:, :, 0, :]), remap(r3[:, :, 1, :]), 0)
            res = r3.reshape(res.shape)
        out[r0:r1] = res
    return out


def dense(formula, aliases, args, dtype="float64"):
    """Dense ``(nrow, ncol, dimF)`` values of the reduction's formula (small cases only); used by the
    tests to decide whether an index mismatch is a genuine near-tie."""
    red = parse(formula, aliases)
    tagI = red.tagI
    nrow = ncol = None
    for v in red.vars().values():
        if v.cat == 2:
            continue
        n = np.asarray(args[v.ind]).reshape(-1, v.dim).shape[0]
        if v.cat == tagI:
            nrow = n
        else:
            ncol = n
    env = _Env(args, np.dtype(dtype), tagI, slice(0, nrow))
    return np.broadcast_to(red.F.ev(env), (nrow, ncol, red.F.dim)).copy()


def abs_scale(formula, aliases, args, dtype="float64", sizes=None, chunk_elems=1 << 24):
    """Per-output *condition scale* ``A`` of a reduction, same shape as :func:`genred`'s result: the value against which a
    relative error of the summation is meaningful.  ``Sum``: ``A_i = sum_j |F_ij|`` (equals ``|out_i|`` when all terms
    have one sign; when terms cancel, an error of ``rtol * A_i`` is what a ``rtol``-accurate evaluation of every term
    can guarantee).  ``LSE``: ``A = (|m|, sum_j |exp(F-m) G|)``.  Extremum-type reductions: ``A = |value|`` (index slots: 0).
    Evaluated in ``dtype`` (float64 by default) -- test infrastructure for the element-wise parity gate."""
    red = parse(formula, aliases)
    dtype = np.dtype(dtype)
    tagI, F, G, K, kind = red.tagI, red.F, red.G, red.K, red.kind
    nrow = ncol = None
    if sizes is not None:
        nrow, ncol = (sizes[0], sizes[1]) if tagI == 0 else (sizes[1], sizes[0])
    for v in red.vars().values():
        if v.cat == 2:
            continue
        n = np.asarray(args[v.ind]).reshape(-1, v.dim).shape[0]
        if v.cat == tagI:
            nrow = n
        else:
            ncol = n
    ref = genred(formula, aliases, args, dtype, chunk_elems=chunk_elems, sizes=sizes)
    if kind not in ("Sum", "LSE") or nrow == 0 or ncol == 0:
        return np.abs(ref)
    d = F.dim
    out = np.zeros_like(ref)
    step = max(1, chunk_elems // max(1, ncol * max(d, 1)))
    for r0 in range(0, nrow, step):
        rows = slice(r0, min(nrow, r0 + step))
        nr = rows.stop - rows.start
        env = _Env(args, dtype, tagI, rows)
        Fv = np.broadcast_to(F.ev(env), (nr, ncol, d))
        if kind == "Sum":
            out[rows] = np.abs(Fv).sum(axis=1, dtype=np.float64)
        else:
            Gv = np.broadcast_to(G.ev(env), (nr, ncol, G.dim))
            m = Fv.max(axis=1)
            out[rows, 0:1] = np.abs(m)
            out[rows, 1:] = np.abs(np.exp(Fv - m[:, None, :]) * Gv).sum(axis=1, dtype=np.float64)
    return out
