"""``LazyTensor``: symbolic (M, N, D) tensors whose reductions run on the sm_100a map-reduce engine.

API of ``pykeops.torch.LazyTensor`` (``pykeops/pykeops/common/lazy_tensor.py`` +
``pykeops/pykeops/torch/lazytensor/LazyTensor.py``), written from scratch: a LazyTensor only
carries a *formula string* in the reference's grammar plus the list of data tensors it refers to;
arithmetic builds bigger strings; a reduction hands the string to :class:`Genred`
(``lazy_tensor.py:648-777``).  Data tensors appear in formulas as ``Var(<id>,dim,cat)`` placeholders
that are renumbered to consecutive indices right before the reduction (``lazy_tensor.py:274-328``).

Supported surface: constructors from tensors / floats / ints / lists / ``(ind, dim, cat)`` triplets,
``Vi / Vj / Pm``, element-wise arithmetic and math functions, ``sum(-1)``, ``|``, ``@``,
concatenation / extraction, all reductions of the engine (sum, logsumexp, sumsoftmaxweight, min,
argmin, min_argmin, max, argmax, max_argmax, Kmin, argKmin, Kmin_argKmin), ``solve`` and batch
dimensions, block-sparse ``ranges`` (helpers in :mod:`keops_b200.torch.cluster`), tensor contractions (``keops_tensordot``,
``keops_kron``), ``bspline``, ``softdtw_sqdist``, and complex numbers (``ComplexLazyTensor``: complex tensors are viewed as
interleaved (re, im) pairs, ``lazy_tensor.py:2669-2823``).  Not carried over: float16.
"""
from __future__ import annotations

import copy
import math
import re

import torch

from .genred import Genred
from .kernel_solve import KernelSolve

_INIT_KWARGS = ("dtype_acc", "use_double_acc", "sum_scheme", "enable_chunks", "use_fast_math", "optional_flags")


def _is_lt(x):
    return isinstance(x, LazyTensor)


def _is_scalar(x, val=None):
    ok = isinstance(x, (int, float, complex)) and not isinstance(x, bool)
    return ok if val is None else (ok and x == val)


def _detect_complex(x):
    """``torchtools.detect_complex`` (``pykeops/pykeops/torch/utils.py:99-105``)"""
    if isinstance(x, list):
        return any(type(v) == complex for v in x)
    if isinstance(x, torch.Tensor):
        return torch.is_complex(x)
    return type(x) == complex


class LazyTensor:
    # hooks for the numpy front-end (keops_b200/numpy): which callable classes run the reductions, and how a
    # user-provided array becomes the torch tensor the bookkeeping below works with
    _Genred = Genred
    _KernelSolve = KernelSolve

    @staticmethod
    def _ingest(x):
        return x

    def __init__(self, x=None, axis=None, is_complex=False):
        self.variables = ()
        self.symbolic_variables = ()
        self.formula = None
        self.formula2 = None
        self.ndim = None
        self.axis = None
        self.ni = None
        self.nj = None
        self.batchdims = None
        self._dtype = None
        self.ranges = None
        self.backend = None
        self.is_complex = False
        if x is None:
            return
        x = self._ingest(x)
        if is_complex or _detect_complex(x):
            # complex data are stored as interleaved (re, im) pairs along the last dimension (lazy_tensor.py:2672-2687)
            self.is_complex = True
            if type(x) == complex:
                x = [x]
            if isinstance(x, list):
                x = [c for v in x for c in (complex(v).real, complex(v).imag)]
            elif isinstance(x, torch.Tensor) and torch.is_complex(x):
                x = torch.view_as_real(x).reshape(tuple(x.shape[:-1]) + (2 * x.shape[-1],)) if x.dim() else torch.view_as_real(x)
        if isinstance(x, tuple):
            if len(x) != 3 or not all(isinstance(v, int) for v in x):
                raise ValueError("LazyTensors(tuple) is only valid if tuple = (ind,dim,cat) is a triplet of integers.")
            if axis is not None:
                raise ValueError("'axis' parameter should not be given when 'x' is of the form (ind,dim,cat).")
            self.symbolic_variables = (x,)
            self.ind, self.ndim, self.axis = x
            self.formula = "VarSymb(%d,%d,%d)" % x
            return
        if isinstance(x, bool):
            raise TypeError("LazyTensors cannot be built from booleans.")
        if isinstance(x, int):
            self.formula = "IntCst(%d)" % x
            self.ndim, self.axis = 1, 2
            return
        if isinstance(x, float):
            x = [x]
        if isinstance(x, list):
            if axis is not None and axis != 2:
                raise ValueError("Lists of numbers are handled as Parameter variables, with an optional 'axis' argument that is equal to 2.")
            self.variables = (x,)
            self.ndim, self.axis = len(x), 2
            self.formula = "Var(%d,%d,2)" % (id(x), self.ndim)
            return
        if not isinstance(x, torch.Tensor):
            raise TypeError("LazyTensors should be built from torch tensors, float/integer numbers, lists of floats "
                            "or 3-uples of integers. Received: {}".format(type(x)))
        if x.dtype not in (torch.float32, torch.float64):
            raise ValueError("[KeOps-B200] LazyTensors support float32 and float64 tensors only (got %s)." % x.dtype)
        self._dtype = str(x.dtype).split(".")[1]
        if x.dim() == 0:
            x = x.view(1)
        if x.dim() >= 3:
            if axis is not None:
                raise ValueError("'axis' parameter should not be given when 'x' is a 3D tensor.")
            if x.dim() > 3:
                self.batchdims = tuple(x.shape[:-3])
            if x.shape[-3] == 1:
                if x.shape[-2] == 1:
                    x, axis = x.squeeze(-2).squeeze(-2), 2
                else:
                    x, axis = x.squeeze(-3), 1
            elif x.shape[-2] == 1:
                x, axis = x.squeeze(-2), 0
            else:
                raise ValueError("If 'x' is a 3D+ tensor, its shape should be one of (..,M,1,D), (..,1,N,D) or (..,1,1,D).")
        if x.dim() >= 2 and axis != 2:
            if axis not in (0, 1):
                raise ValueError("When 'x' is encoded as a 2D array, LazyTensor expects an explicit 'axis' value in {0,1}.")
            self.variables = (x,)
            self.ndim, self.axis = x.shape[-1], axis
            self.formula = "Var(%d,%d,%d)" % (id(x), self.ndim, axis)
            if axis == 0:
                self.ni = x.shape[-2]
            else:
                self.nj = x.shape[-2]
            if x.dim() > 2 and self.batchdims is None:
                self.batchdims = tuple(x.shape[:-2])
        else:
            if axis not in (None, 2):
                raise ValueError("When 'x' is encoded as a 1D or 0D array, 'axis' must be None or 2 (= Parameter variable).")
            self.variables = (x,)
            self.ndim, self.axis = x.shape[-1], 2
            self.formula = "Var(%d,%d,2)" % (id(x), self.ndim)
            if x.dim() > 1 and self.batchdims is None:
                self.batchdims = tuple(x.shape[:-1])

    # ------------------------------------------------------------------------------------------
    # bookkeeping
    # ------------------------------------------------------------------------------------------
    @property
    def nbatchdims(self):
        return 0 if self.batchdims is None else len(self.batchdims)

    @property
    def dtype(self):
        return self._dtype

    @property
    def _shape(self):
        bd = () if self.batchdims is None else tuple(self.batchdims)
        ni = 1 if self.ni is None else self.ni
        nj = 1 if self.nj is None else self.nj
        nd = 1 if self.ndim is None else self.ndim
        return bd + (ni, nj, nd)

    @property
    def shape(self):
        """(batch..., M, N, D), the trailing D being dropped when it is 1 (``lazy_tensor.py:1003-1009``)."""
        s = self._shape
        if self.is_complex:
            s = s[:-1] + (s[-1] // 2,)
        return s[:-1] if s[-1] == 1 else s

    def dim(self):
        """Number of dimensions *including* the trailing vector dimension (``lazy_tensor.py:1011-1015``)."""
        return len(self._shape)

    def new_variable_index(self):
        sv = self.symbolic_variables
        return 0 if len(sv) == 0 else 1 + max(v[0] for v in sv)

    def _copy_meta(self):
        r = type(self)()
        for k in ("variables", "symbolic_variables", "ni", "nj", "batchdims", "_dtype", "ranges", "backend"):
            setattr(r, k, getattr(self, k))
        return r

    def _join(self, other):
        """Merge the bookkeeping of two operands (``lazy_tensor.py:355-430``)."""
        r = type(self)()
        seen, vs = set(), []
        for v in self.variables + other.variables:
            if id(v) not in seen:
                seen.add(id(v))
                vs.append(v)
        r.variables = tuple(vs)
        r.symbolic_variables = tuple(dict.fromkeys(self.symbolic_variables + other.symbolic_variables))
        for k in ("ni", "nj"):
            a, b = getattr(self, k), getattr(other, k)
            if a is not None and b is not None and a != b:
                raise ValueError("Incompatible sizes along the '%s' dimension: %d vs %d." % (k[1], a, b))
            setattr(r, k, a if a is not None else b)
        if self._dtype is not None and other._dtype is not None and self._dtype != other._dtype:
            raise ValueError("Incompatible dtypes: %s and %s." % (self._dtype, other._dtype))
        r._dtype = self._dtype if self._dtype is not None else other._dtype
        if self.batchdims is None:
            r.batchdims = other.batchdims
        elif other.batchdims is None:
            r.batchdims = self.batchdims
        else:
            if len(self.batchdims) != len(other.batchdims):
                raise ValueError("Incompatible number of batch dimensions: {} and {}.".format(self.batchdims, other.batchdims))
            r.batchdims = tuple(torch.broadcast_shapes(self.batchdims, other.batchdims))
        r.ranges = self.ranges if self.ranges is not None else other.ranges
        r.backend = self.backend if self.backend is not None else other.backend
        return r

    def _wrap(self, x):
        return x if _is_lt(x) else type(self)(x)

    def unary(self, operation, dimres=None, opt_arg=None, opt_arg2=None, is_complex=False):
        r = self._copy_meta()
        r.is_complex = is_complex
        r.ndim = self.ndim if dimres is None else dimres
        extra = "".join(",%s" % a for a in (opt_arg, opt_arg2) if a is not None)
        r.formula = "%s(%s%s)" % (operation, self.formula, extra)
        return r

    def binary(self, other, operation, is_operator=False, dimres=None, dimcheck="sameor1", opt_arg=None, opt_pos="last",
               is_complex=False, rversion=False):
        other = self._wrap(other)
        if rversion:                     # "other <op> self" (lazy_tensor.py:540-545)
            return other.binary(self, operation, is_operator=is_operator, dimres=dimres, dimcheck=dimcheck, opt_arg=opt_arg,
                                opt_pos=opt_pos, is_complex=is_complex)
        if dimres is None:
            dimres = max(self.ndim, other.ndim)
        if dimcheck == "same" and self.ndim != other.ndim:
            raise ValueError("Operation %s expects inputs of the same dimension. Received %d and %d." % (operation, self.ndim, other.ndim))
        if dimcheck == "sameor1" and self.ndim != other.ndim and self.ndim != 1 and other.ndim != 1:
            raise ValueError("Operation %s expects inputs of the same dimension or dimension 1. Received %d and %d."
                             % (operation, self.ndim, other.ndim))
        r = self._join(other)
        r.is_complex = is_complex
        r.ndim = dimres
        if is_operator:
            r.formula = "(%s %s %s)" % (self.formula, operation, other.formula)
        elif opt_arg is not None:
            if opt_pos == "last":
                r.formula = "%s(%s, %s, %s)" % (operation, self.formula, other.formula, opt_arg)
            else:                       # "middle" (lazy_tensor.py:557-560)
                r.formula = "%s(%s, %s, %s)" % (operation, self.formula, opt_arg, other.formula)
        else:
            r.formula = "%s(%s, %s)" % (operation, self.formula, other.formula)
        return r

    def ternary(self, other1, other2, operation, dimres=None):
        other1, other2 = self._wrap(other1), self._wrap(other2)
        if dimres is None:
            dimres = max(self.ndim, other1.ndim, other2.ndim)
        r = self._join(other1)._join_keep(other2)
        r.ndim = dimres
        r.formula = "%s(%s, %s, %s)" % (operation, self.formula, other1.formula, other2.formula)
        return r

    def _join_keep(self, other):
        return self._join(other)

    # ------------------------------------------------------------------------------------------
    # arithmetic
    # ------------------------------------------------------------------------------------------
    # The four arithmetic operators dispatch on complex operands the way the reference does (lazy_tensor.py:1023-1175,
    # 2775-2807): a real operand meeting a complex one is promoted (real2complex) -- except in a product, where a real
    # scalar factor becomes ComplexRealScal; python complex numbers become complex parameter LazyTensors.
    def _addop(self, other, **kw):
        if self.is_complex:
            other = self._wrap(other)
            if not other.is_complex:
                return self._addop(other.real2complex(), **kw)
            return self.binary(other, "ComplexAdd", is_complex=True, dimcheck=None, **kw)
        return self.binary(other, "+", is_operator=True, **kw)

    def _subop(self, other, **kw):
        if self.is_complex:
            other = self._wrap(other)
            if not other.is_complex:
                return self._subop(other.real2complex(), **kw)
            return self.binary(other, "ComplexSubtract", is_complex=True, dimcheck=None, **kw)
        return self.binary(other, "-", is_operator=True, **kw)

    def _mulop(self, other, **kw):
        if self.is_complex:
            other = self._wrap(other)
            if other._shape[-1] == 1:
                return other.binary(self, "ComplexRealScal", is_complex=True, **kw)
            if not other.is_complex:
                return self._mulop(other.real2complex(), **kw)
            return self.binary(other, "ComplexMult", is_complex=True, dimcheck=None, **kw)
        return self.binary(other, "*", is_operator=True, **kw)

    def _divop(self, other, **kw):
        if self.is_complex:
            other = self._wrap(other)
            if not other.is_complex:
                return self._divop(other.real2complex(), **kw)
            return self.binary(other, "ComplexDivide", is_complex=True, dimcheck=None, **kw)
        return self.binary(other, "/", is_operator=True, **kw)

    def _other_is_complex(self, other):
        return _is_lt(other) and other.is_complex

    def __add__(self, other):
        if _is_scalar(other, 0):
            return self
        if self._other_is_complex(other) and not self.is_complex:
            return self.real2complex()._addop(other)
        return self._addop(other)

    def __radd__(self, other):
        if _is_scalar(other, 0):
            return self
        return self._addop(other, rversion=True)

    def __sub__(self, other):
        if _is_scalar(other, 0):
            return self
        if self._other_is_complex(other) and not self.is_complex:
            return self.real2complex()._subop(other)
        return self._subop(other)

    def __rsub__(self, other):
        if _is_scalar(other, 0):
            return self.unary("Minus")
        return self._subop(other, rversion=True)

    def __mul__(self, other):
        if _is_scalar(other, 0):
            return 0
        if _is_scalar(other, 1):
            return self
        if _is_scalar(other, -1):
            return self.unary("Minus")
        if self._other_is_complex(other) and not self.is_complex:
            return other._mulop(self)
        if _detect_complex(other) and not self.is_complex:
            return type(self)(other)._mulop(self)
        return self._mulop(other)

    def __rmul__(self, other):
        if _is_scalar(other, 0):
            return 0
        if _is_scalar(other, 1):
            return self
        if _is_scalar(other, -1):
            return self.unary("Minus")
        if _detect_complex(other) and not self.is_complex:
            return self.real2complex()._mulop(type(self)(other))
        return self._mulop(other, rversion=True)

    def __truediv__(self, other):
        if _is_scalar(other, 1):
            return self
        if self._other_is_complex(other) and not self.is_complex:
            return self.real2complex()._divop(other)
        return self._divop(other)

    def __rtruediv__(self, other):
        if _is_scalar(other, 0):
            return 0
        if _is_scalar(other, 1):
            return self.unary("Inv")
        return self._divop(other, rversion=True)

    def __neg__(self):
        return self.unary("Minus")

    def __abs__(self):
        if self.is_complex:
            return self.unary("ComplexAbs", dimres=self.ndim // 2)
        return self.unary("Abs")

    # ---- complex numbers (lazy_tensor.py:1357-1400, 2669-2823) ----------------------------------
    def _need_complex(self, what):
        if not self.is_complex:
            raise ValueError("%s is only defined for complex LazyTensors." % what)

    def _need_real(self, what):
        if self.is_complex:
            raise ValueError("%s cannot be applied to a complex LazyTensor." % what)

    @property
    def real(self):
        self._need_complex("real")
        return self.unary("ComplexReal", dimres=self.ndim // 2)

    @property
    def imag(self):
        self._need_complex("imag")
        return self.unary("ComplexImag", dimres=self.ndim // 2)

    def angle(self):
        self._need_complex("angle")
        return self.unary("ComplexAngle", dimres=self.ndim // 2)

    def conj(self):
        self._need_complex("conj")
        return self.unary("Conj", dimres=self.ndim, is_complex=True)

    def real2complex(self):
        self._need_real("real2complex")
        return self.unary("Real2Complex", dimres=2 * self.ndim, is_complex=True)

    def imag2complex(self):
        self._need_real("imag2complex")
        return self.unary("Imag2Complex", dimres=2 * self.ndim, is_complex=True)

    def exp1j(self):
        self._need_real("exp1j")
        return self.unary("ComplexExp1j", dimres=2 * self.ndim, is_complex=True)

    def abs(self):
        return abs(self)

    def __or__(self, other):
        return self.binary(other, "|", is_operator=True, dimres=1, dimcheck="same")

    def __ror__(self, other):
        return self._wrap(other).binary(self, "|", is_operator=True, dimres=1, dimcheck="same")

    def __pow__(self, other):
        if isinstance(other, int):
            return self.unary("Square") if other == 2 else self.unary("Pow", opt_arg=other)
        if isinstance(other, float):
            if other == 0.5:
                return self.unary("Sqrt")
            if other == -0.5:
                return self.unary("Rsqrt")
            other = type(self)(other)
        if _is_lt(other):
            if other.ndim in (1, self.ndim):
                return self.binary(other, "Powf", dimcheck=None)
            raise ValueError("Incompatible dimensions for the LazyTensor and its exponent: %d and %d." % (self.ndim, other.ndim))
        raise ValueError("The exponent should be an integer, a float number or a LazyTensor.")

    def power(self, other):
        return self ** other

    def __lt__(self, other):
        return self.binary(other, "LessThan")

    def __gt__(self, other):
        return self._wrap(other).binary(self, "LessThan")

    def __le__(self, other):
        return self.binary(other, "LessOrEqual")

    def __ge__(self, other):
        return self._wrap(other).binary(self, "LessOrEqual")

    def __eq__(self, other):
        return self.binary(other, "Equal")

    def __ne__(self, other):
        return self.binary(other, "NotEqual")

    __hash__ = object.__hash__

    # ---- element-wise math -------------------------------------------------------------------
    def exp(self): return self.unary("ComplexExp", dimres=self.ndim, is_complex=True) if self.is_complex else self.unary("Exp")
    def log(self): return self.unary("Log")
    def xlogx(self): return self.unary("XLogX")
    def cos(self): return self.unary("Cos")
    def sin(self): return self.unary("Sin")
    def sinxdivx(self): return self.unary("SinXDivX")
    def asin(self): return self.unary("Asin")
    def acos(self): return self.unary("Acos")
    def atan(self): return self.unary("Atan")
    def atan2(self, other): return self.binary(other, "Atan2", dimcheck="sameor1")
    def sqrt(self): return self.unary("Sqrt")
    def rsqrt(self): return self.unary("Rsqrt")
    def square(self): return self.unary("Square")
    def sign(self): return self.unary("Sign")
    def step(self): return self.unary("Step")
    def relu(self): return self.unary("ReLU")
    def floor(self): return self.unary("Floor")
    def round(self, other=0): return self.unary("Round", opt_arg=other)
    def mod(self, modulus, offset=0): return self.ternary(modulus, offset, "Mod")
    def ifelse(self, other1, other2): return self.ternary(other1, other2, "IfElse")

    def clamp(self, other1, other2):
        if isinstance(other1, int) and isinstance(other2, int):
            return self.unary("ClampInt", opt_arg=other1, opt_arg2=other2)
        return self.ternary(other1, other2, "Clamp")

    def sqnorm2(self): return self.unary("SqNorm2", dimres=1)
    def norm2(self): return self.unary("Norm2", dimres=1)

    def norm(self, dim):
        if dim not in (-1, len(self._shape) - 1):
            raise ValueError("KeOps only supports norms over the last dimension.")
        return self.norm2()

    def normalize(self): return self.unary("Normalize")
    def sqdist(self, other): return self.binary(other, "SqDist", dimres=1)

    def weightedsqnorm(self, other):
        other = self._wrap(other)
        if other.ndim not in (1, self.ndim, self.ndim ** 2):
            raise ValueError("Squared norm weights should be of size 1 (scalar), D (diagonal) or D^2 (full symmetric tensor), "
                             "but received {} with D={}.".format(other.ndim, self.ndim))
        return other.binary(self, "WeightedSqNorm", dimres=1, dimcheck=None)

    def weightedsqdist(self, g, s):
        return (self - g).weightedsqnorm(s)

    def elem(self, i):
        if not isinstance(i, int) or not 0 <= i < self.ndim:
            raise ValueError("Elem: index must be an integer in [0, dim).")
        return self.unary("Elem", dimres=1, opt_arg=i)

    def extract(self, i, d):
        if not (isinstance(i, int) and isinstance(d, int)) or i < 0 or i + d > self.ndim:
            raise ValueError("Extract: slice is out of bounds.")
        return self.unary("Extract", dimres=d, opt_arg=i, opt_arg2=d)

    def __getitem__(self, key):
        if isinstance(key, tuple):
            if len(key) == len(self._shape) and all(k == slice(None) for k in key[:-1]):
                key = key[-1]
            else:
                raise ValueError("LazyTensors only support indexing with respect to their last dimension.")
        if isinstance(key, slice):
            if key.step not in (None, 1):
                raise ValueError("LazyTensors do not support sliced indexing with stepsizes > 1.")
            start = 0 if key.start is None else key.start
            stop = self.ndim if key.stop is None else key.stop
            if start < 0:
                start += self.ndim
            if stop < 0:
                stop += self.ndim
            return self.extract(start, stop - start)
        if isinstance(key, int):
            return self.elem(key + self.ndim if key < 0 else key)
        raise ValueError("LazyTensors only support indexing with integers and vanilla python slices.")

    def one_hot(self, D):
        if self.ndim != 1:
            raise ValueError("One-hot encoding expects scalar input.")
        return self.unary("OneHot", dimres=D, opt_arg=D)

    def concat(self, other):
        other = self._wrap(other)
        return self.binary(other, "Concat", dimres=self.ndim + other.ndim, dimcheck=None)

    @staticmethod
    def concatenate(tuple_of_lt, axis=-1):
        if not isinstance(tuple_of_lt, (tuple, list)) or len(tuple_of_lt) == 0:
            raise ValueError("concatenate expects a non-empty tuple of LazyTensors.")
        if axis != -1:
            raise ValueError("LazyTensor only supports concatenation along the last axis.")
        r = tuple_of_lt[0]
        for t in tuple_of_lt[1:]:
            r = r.concat(t)
        return r

    cat = concatenate

    def matvecmult(self, other):
        other = self._wrap(other)
        return self.binary(other, "MatVecMult", dimres=self.ndim // other.ndim, dimcheck=None)

    def vecmatmult(self, other):
        other = self._wrap(other)
        return self.binary(other, "VecMatMult", dimres=other.ndim // self.ndim, dimcheck=None)

    def tensorprod(self, other):
        other = self._wrap(other)
        return self.binary(other, "TensorProd", dimres=self.ndim * other.ndim, dimcheck=None)

    def trace_operator(self, var):
        """Symbolic trace of ``self``, a linear function of the symbolic variable ``var`` (``lazy_tensor.py:2007-2037``).  The
        variable disappears from the result: it leaves the list of symbolic variables and its index is made negative so that
        it is never mistaken for an argument."""
        if not _is_lt(var) or len(var.symbolic_variables) != 1 or not hasattr(var, "ind"):
            raise ValueError("trace_operator expects a symbolic variable, e.g. Vi(2, 3).")
        res = self.binary(var, "TraceOperator", dimres=1, dimcheck="same")
        res.symbolic_variables = tuple(v for v in res.symbolic_variables if v not in var.symbolic_variables)
        if var.ind >= 0:
            res.formula = res.formula.replace(var.formula, "VarSymb(-%d-1,%d,%d)" % (var.ind, var.ndim, var.axis))
        return res

    def bspline(self, x, k=0):
        """Vector of B-spline basis functions of order ``k`` for the knots ``self``, evaluated at the scalar ``x``
        (``lazy_tensor.py:1706-1714``)."""
        x = self._wrap(x)
        if x.ndim != 1:
            raise ValueError("Operation BSpline expects a vector and a scalar input (of dimension 1). Received %d and %d."
                             % (self.ndim, x.ndim))
        return self.binary(x, "BSpline", dimres=self.ndim - k - 1, dimcheck=None, opt_arg="%d" % k)

    def softdtw_sqdist(self, other, gamma):
        """Soft-DTW discrepancy between the vectors ``self`` and ``other`` with smoothness ``gamma``, squared difference
        as the ground cost (``lazy_tensor.py:1599-1612``)."""
        return self.ternary(other, gamma, "SoftDTW_SqDist", dimres=1)

    def keops_tensordot(self, other, dimfa, dimfb, contfa, contfb, *args):
        if self.is_complex or (_is_lt(other) and other.is_complex):
            raise ValueError("keops_tensordot is not implemented for complex LazyTensors")
        return self._keops_tensordot(other, dimfa, dimfb, contfa, contfb, *args)

    def _keops_tensordot(self, other, dimfa, dimfb, contfa, contfb, *args):
        """Tensor contraction over paired axes of the internal dimension, viewed as row-major tensors of shapes ``dimfa``
        and ``dimfb``; an optional last argument permutes the kept axes (``lazy_tensor.py:1816-1849``)."""
        def seq(v):
            return "[" + ",".join(str(int(k)) for k in ((v,) if isinstance(v, int) else v)) + "]"
        opt_arg = ", ".join(seq(v) for v in (dimfa, dimfb, contfa, contfb) + args)
        da, db = [int(k) for k in dimfa], [int(k) for k in dimfb]
        dimres = 1
        for k in da + db:
            dimres *= k
        for c in ((contfa,) if isinstance(contfa, int) else contfa):
            dimres //= da[int(c)] ** 2
        return self.binary(other, "TensorDot", dimres=dimres, dimcheck=None, opt_arg=opt_arg)

    def keops_kron(self, other, dimfself, dimfother):
        """Kronecker product (``numpy.kron``) of the internal dimensions viewed as tensors of shapes ``dimfself`` and
        ``dimfother`` with the same number of axes (``lazy_tensor.py:1851-1907``, ``maths/Kron.py``)."""
        if isinstance(dimfself, int) or isinstance(dimfother, int):
            raise ValueError("For keops_Kronecker, dimfself and dimfother must be tuple/list of int")
        opt_arg = ", ".join("[" + ",".join(str(int(k)) for k in v) + "]" for v in (dimfself, dimfother))
        return self.binary(other, "Kron", dimres=self.ndim * self._wrap(other).ndim, dimcheck=None, opt_arg=opt_arg)

    # ------------------------------------------------------------------------------------------
    # reductions
    # ------------------------------------------------------------------------------------------
    def _fixvariables(self):
        """Give final consecutive labels to data variables (after the symbolic ones) and pad batch
        dimensions (``lazy_tensor.py:274-328``)."""
        f1 = self.formula
        f2 = self.formula2 or ""
        device, dtype = None, None
        for v in self.variables:
            if isinstance(v, torch.Tensor):
                device, dtype = v.device, v.dtype
                break
        i = self.new_variable_index()
        newvars = []
        for v in self.variables:
            tag = "Var(%d," % id(v)
            if tag not in f1 + f2:
                continue
            f1 = f1.replace(tag, "Var(%d," % i)
            f2 = f2.replace(tag, "Var(%d," % i)
            if isinstance(v, list):
                if device is None:
                    raise ValueError("[KeOps-B200] a formula made of python numbers only has no device to run on.")
                v = torch.tensor(v, dtype=dtype, device=device)
            cat = int(re.search(r"Var\(%d,\d+,([012])\)" % i, f1 + f2).group(1))
            want = self.nbatchdims + 1 + (1 if cat in (0, 1) else 0)
            if v.dim() < want:
                v = v.view((1,) * (want - v.dim()) + tuple(v.shape))
            newvars.append(v)
            i += 1
        self.formula = f1.replace("VarSymb(", "Var(")
        self.formula2 = f2.replace("VarSymb(", "Var(") or None
        self.variables = tuple(newvars)

    @staticmethod
    def _separate_kwargs(kwargs):
        init = {k: v for k, v in kwargs.items() if k in _INIT_KWARGS}
        call = {k: v for k, v in kwargs.items() if k not in _INIT_KWARGS}
        init.pop("optional_flags", None)
        return init, call

    def reduction(self, reduction_op, other=None, opt_arg=None, axis=None, dim=None, call=True, **kwargs):
        if axis is None:
            axis = dim
        if axis is None:
            raise ValueError("Reductions must be called with an 'axis' (or 'dim') argument.")
        if axis < 0:
            axis += self.nbatchdims + 3      # a LazyTensor is a virtual (batch..., M, N, D) array
        if axis - self.nbatchdims not in (0, 1):
            raise ValueError("Reductions must be called with 'axis' (or 'dim') equal to the number of batch dimensions + 0 or 1.")
        if other is None:
            r = self._copy_meta()
        else:
            o = self._wrap(other)        # wrap ONCE: a fresh wrap of a float / array creates a new variable identity
            r = self._join(o)
            r.formula2 = o.formula
        r.formula = self.formula
        r.ndim = self.ndim
        r.is_complex = self.is_complex
        r.reduction_op = reduction_op
        r.axis = axis - self.nbatchdims
        r.opt_arg = opt_arg
        init, r.kwargs = self._separate_kwargs(kwargs)
        r._init_kwargs = init
        if r._dtype is not None:
            r._fixvariables()
            r.callfun = self._Genred(r.formula, [], reduction_op=reduction_op, axis=r.axis, opt_arg=opt_arg,
                               formula2=r.formula2, **init)
        if call and len(r.symbolic_variables) == 0 and r._dtype is not None:
            return r()
        return r

    def solve(self, other, var=None, call=True, **kwargs):
        """``K.solve(b, alpha=...)``: solve ``(alpha Id + K) a = b`` by conjugate gradients
        (``lazy_tensor.py:779-870``)."""
        if not _is_lt(other):
            other = type(self)(other, axis=0)
        if len(other.symbolic_variables) == 0 and len(self.symbolic_variables) != 0:
            raise ValueError("If 'self' has symbolic variables, so should 'other'.")
        axis = 1 - other.axis
        if var is None:
            varindex = self.new_variable_index()
            var = type(self)((varindex, other.ndim, axis))
            r = self * var
        else:
            varindex = var.symbolic_variables[0][0]
            r = self._copy_meta()
            r.formula = self.formula
            r.ndim = self.ndim
        r.formula2 = None
        r.reduction_op = "Solve"
        r.varindex = varindex
        r.varformula = var.formula.replace("VarSymb", "Var")
        r.other = other
        r.axis = axis
        init, r.kwargs = self._separate_kwargs(kwargs)
        r._init_kwargs = init
        if r._dtype is None:
            r._dtype = other._dtype
        if r._dtype is not None:
            r._fixvariables()
            r.callfun = self._KernelSolve(r.formula, [], r.varformula, r.axis, **init)
        if call and len(other.symbolic_variables) == 0 and r._dtype is not None:
            return r()
        return r

    def __call__(self, *args, **kwargs):
        if not hasattr(self, "reduction_op"):
            raise ValueError("A LazyTensor object may be called only if it corresponds to the output of a reduction operation or solve operation.")
        self.kwargs.update(kwargs)
        if self.ranges is not None and "ranges" not in self.kwargs:
            self.kwargs["ranges"] = self.ranges
        if self.backend is not None and "backend" not in self.kwargs:
            self.kwargs["backend"] = self.backend
        if self._dtype is None:
            self._dtype = str(args[0].dtype).split(".")[-1]
            self._fixvariables()
            if self.reduction_op == "Solve":
                self.callfun = self._KernelSolve(self.formula, [], self.varformula, self.axis, **self._init_kwargs)
            else:
                self.callfun = self._Genred(self.formula, [], reduction_op=self.reduction_op, axis=self.axis,
                                      opt_arg=self.opt_arg, formula2=self.formula2, **self._init_kwargs)
        if self.reduction_op == "Solve" and len(self.other.symbolic_variables) == 0:
            if args != ():
                raise ValueError("no input required")
            args = (self.other.variables[0],)
        res = self.callfun(*args, *self.variables, **self.kwargs)
        return self._view_as_complex(res) if self.is_complex else res

    @staticmethod
    def _view_as_complex(res):
        """interleaved (re, im) pairs -> complex tensor (``pykeops/pykeops/torch/utils.py:107-113``)"""
        if isinstance(res, torch.Tensor) and not torch.is_complex(res):
            return torch.view_as_complex(res.reshape(tuple(res.shape[:-1]) + (res.shape[-1] // 2, 2)).contiguous())
        return res

    def __str__(self):
        tail = ""
        if hasattr(self, "reduction_op"):
            tail = ", reduction=%s, axis=%s" % (self.reduction_op, self.axis)
        return "KeOps-B200 LazyTensor\n    formula: %s\n    shape: %s%s" % (self.formula, self.shape, tail)

    __repr__ = __str__

    # -- the reductions ------------------------------------------------------------------------
    def sum(self, axis=-1, dim=None, **kwargs):
        if dim is not None:
            axis = dim
        if axis in (-1, 2 + self.nbatchdims):
            if self.is_complex:
                return self.unary("ComplexSum", dimres=2, is_complex=True)
            return self.unary("Sum", dimres=1)
        return self.reduction("Sum", axis=axis, **kwargs)

    def sum_reduction(self, axis=None, dim=None, **kwargs):
        return self.reduction("Sum", axis=axis, dim=dim, **kwargs)

    def logsumexp(self, axis=None, dim=None, weight=None, **kwargs):
        if weight is None:
            return self.reduction("LogSumExp", axis=axis, dim=dim, **kwargs)
        return self.reduction("LogSumExp", other=weight, axis=axis, dim=dim, **kwargs)

    def logsumexp_reduction(self, **kwargs):
        return self.logsumexp(**kwargs)

    def sumsoftmaxweight(self, weight, axis=None, dim=None, **kwargs):
        return self.reduction("SumSoftMaxWeight", other=weight, axis=axis, dim=dim, **kwargs)

    def sumsoftmaxweight_reduction(self, **kwargs):
        return self.sumsoftmaxweight(**kwargs)

    def _vec_or_red(self, unary_name, red_name, axis, dim, kwargs, opt_arg=None):
        if dim is not None:
            axis = dim
        if axis in (-1, 2 + self.nbatchdims):
            return self.unary(unary_name, dimres=1)
        return self.reduction(red_name, axis=axis, opt_arg=opt_arg, **kwargs)

    def min(self, axis=-1, dim=None, **kwargs): return self._vec_or_red("Min", "Min", axis, dim, kwargs)
    def max(self, axis=-1, dim=None, **kwargs): return self._vec_or_red("Max", "Max", axis, dim, kwargs)
    def argmin(self, axis=-1, dim=None, **kwargs): return self._vec_or_red("ArgMin", "ArgMin", axis, dim, kwargs)
    def argmax(self, axis=-1, dim=None, **kwargs): return self._vec_or_red("ArgMax", "ArgMax", axis, dim, kwargs)
    def min_reduction(self, axis=None, dim=None, **kwargs): return self.reduction("Min", axis=axis, dim=dim, **kwargs)
    def max_reduction(self, axis=None, dim=None, **kwargs): return self.reduction("Max", axis=axis, dim=dim, **kwargs)
    def argmin_reduction(self, axis=None, dim=None, **kwargs): return self.reduction("ArgMin", axis=axis, dim=dim, **kwargs)
    def argmax_reduction(self, axis=None, dim=None, **kwargs): return self.reduction("ArgMax", axis=axis, dim=dim, **kwargs)
    def min_argmin(self, axis=None, dim=None, **kwargs): return self.reduction("Min_ArgMin", axis=axis, dim=dim, **kwargs)
    def max_argmax(self, axis=None, dim=None, **kwargs): return self.reduction("Max_ArgMax", axis=axis, dim=dim, **kwargs)
    min_argmin_reduction = min_argmin
    max_argmax_reduction = max_argmax

    def Kmin(self, K, axis=None, dim=None, **kwargs): return self.reduction("KMin", opt_arg=K, axis=axis, dim=dim, **kwargs)
    def argKmin(self, K, axis=None, dim=None, **kwargs): return self.reduction("ArgKMin", opt_arg=K, axis=axis, dim=dim, **kwargs)
    def Kmin_argKmin(self, K, axis=None, dim=None, **kwargs): return self.reduction("KMin_ArgKMin", opt_arg=K, axis=axis, dim=dim, **kwargs)
    Kmin_reduction = Kmin
    argKmin_reduction = argKmin
    Kmin_argKmin_reduction = Kmin_argKmin

    def __matmul__(self, v, **kwargs):
        """``K @ v`` = ``(K * Vj(v)).sum(axis of j)`` (``lazy_tensor.py:2490-2527``)."""
        if self._shape[-1] != 1:
            raise ValueError("The 'K @ v' syntax is only supported for LazyTensors 'K' whose trailing dimension "
                             "is equal to 1. Here, K.shape = {}.".format(self.shape))
        if not _is_lt(v):
            v = self._ingest(v)
        if _is_lt(v):
            Kv = self * v
        else:
            if v.dim() == 1:
                newdims = (1, v.shape[0], 1)
            else:
                newdims = tuple(v.shape[:-2]) + (1,) + tuple(v.shape[-2:])
            Kv = self * type(self)(v.view(newdims))
        out = Kv.sum(Kv.dim() - 2, **kwargs)
        if not _is_lt(v) and v.dim() == 1:
            out = out.reshape(-1)
        return out

    def t(self):
        """Matrix transposition: the i- and j-variables trade places, ``K.t().shape == (..., N, M, D)``
        (``pykeops/pykeops/common/lazy_tensor.py:2529-2601``)."""
        from .cluster import swap_axes
        r = copy.copy(self)
        r.ni, r.nj = self.nj, self.ni
        r.ranges = swap_axes(self.ranges)
        if getattr(r, "axis", None) in (0, 1):
            r.axis = 1 - r.axis

        def swap(f):
            if f is None:
                return None
            f = re.sub(r"(Var|VarSymb)\((\d+),(\d+),0\)", r"\1(\2,\3,i)", f)
            f = re.sub(r"(Var|VarSymb)\((\d+),(\d+),1\)", r"\1(\2,\3,0)", f)
            return re.sub(r"(Var|VarSymb)\((\d+),(\d+),i\)", r"\1(\2,\3,1)", f)

        r.formula, r.formula2 = swap(self.formula), swap(self.formula2)
        r.symbolic_variables = tuple((ind, dim, (1 - cat) if cat in (0, 1) else cat)
                                     for ind, dim, cat in self.symbolic_variables)
        # variables are labelled by id(): the transposed copy needs ids of its own, otherwise K and K.t() could
        # not appear in the same formula (the same tensor would be an i- and a j-variable at once)
        newvars = []
        for v in self.variables:
            if isinstance(v, list):
                newvars.append(v)
                continue
            w = v.view(v.shape)
            newvars.append(w)
            for attr in ("formula", "formula2"):
                f = getattr(r, attr)
                if f is not None:
                    setattr(r, attr, f.replace("Var(%d," % id(v), "Var(%d," % id(w)))
        r.variables = tuple(newvars)
        return r

    @property
    def T(self):
        """Numpy-friendly alias of :meth:`t` (``lazy_tensor.py:2603-2608``)."""
        return self.t()

    def matvec(self, v):
        """``K @ v`` under the name :mod:`scipy.sparse.linalg` expects (``lazy_tensor.py:2610-2631``)."""
        return self @ v

    def rmatvec(self, v):
        """``K.T @ v`` (``lazy_tensor.py:2633-2639``)."""
        return self.T @ v

    def grad(self, other, gradin):
        """Symbolic gradient: ``x.grad(v, e)`` encodes the adjoint of the differential of ``x`` with respect to the
        variable ``v``, applied to ``e`` (``lazy_tensor.py:1918-1936``; formula-level ``Grad``)."""
        return self.binary(gradin, "Grad", dimres=other.ndim, dimcheck="same", opt_arg=other.formula, opt_pos="middle")

    def diff(self, other, diffin):
        """Symbolic differential (forward mode): ``x.diff(v, e)`` encodes the differential of ``x`` with respect to the
        variable ``v``, applied to ``e`` (``lazy_tensor.py:1938-1956``; formula-level ``Diff``)."""
        return self.binary(diffin, "Diff", dimres=self.ndim, dimcheck=None, opt_arg=other.formula, opt_pos="middle")

    def factorize(self, other):
        """``x.factorize(y)`` (``lazy_tensor.py:1958-1973``): common sub-expressions are always shared by this engine."""
        return self.binary(other, "Factorize", dimres=self.ndim, dimcheck=None)

    def auto_factorize(self):
        """``x.auto_factorize()`` (``lazy_tensor.py:1975-1988``): a no-op here, see :meth:`factorize`."""
        return self.unary("AutoFactorize", dimres=self.ndim)

    def grad_matrix(self, other):
        """Symbolic Jacobian of ``x`` with respect to the variable ``other``, flattened to a vector of dimension
        ``x.ndim * other.ndim`` (``lazy_tensor.py:1990-2006``)."""
        return self.unary("GradMatrix", dimres=self.ndim * other.ndim, opt_arg=other.formula)

    def divergence(self, var):
        """Symbolic divergence of ``x`` with respect to the variable ``var`` (``lazy_tensor.py:2038-2052``)."""
        return self.binary(var, "Divergence", dimres=1, dimcheck="same")

    def laplacian(self, var):
        """Symbolic Laplacian of the scalar ``x`` with respect to ``var`` (``lazy_tensor.py:2054-2068``)."""
        return self.binary(var, "Laplacian", dimres=1, dimcheck=None)

    def sinc(self):
        """``sin(pi x) / (pi x)`` (``lazy_tensor.py:1267-1274``)."""
        return (math.pi * self).sinxdivx()


class ComplexLazyTensor(LazyTensor):
    """``ComplexLazyTensor(x)``: a LazyTensor forced to be complex (``lazy_tensor.py:2669-2687``).  Complex-ness is a flag of
    :class:`LazyTensor` here, so this constructor simply returns ``LazyTensor(x, axis, is_complex=True)``."""

    def __new__(cls, x=None, axis=None):
        return LazyTensor(x, axis, is_complex=True)


def Vi(x_or_ind, dim=None):
    """``Vi(x)`` for an (M, D) tensor, or the symbolic ``Vi(ind, dim)``."""
    return LazyTensor((x_or_ind, dim, 0)) if isinstance(x_or_ind, int) else LazyTensor(x_or_ind, axis=0)


def Vj(x_or_ind, dim=None):
    return LazyTensor((x_or_ind, dim, 1)) if isinstance(x_or_ind, int) else LazyTensor(x_or_ind, axis=1)


def Pm(x_or_ind, dim=None):
    if isinstance(x_or_ind, int) and dim is not None:
        return LazyTensor((x_or_ind, dim, 2))
    return LazyTensor(x_or_ind, axis=2) if isinstance(x_or_ind, torch.Tensor) else LazyTensor(x_or_ind)
