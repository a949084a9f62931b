"""Scalar DAG: the code generator's working IR.

The vector formula IR (:mod:`keops_b200.formulas.ir`) is lowered component by component into a
hash-consed DAG of scalar operations.  This is the replacement for the reference's string-pasting
toolbox (``keopscore/keopscore/utils/code_gen_utils.py``) and for its ``Factorize`` pass: common
sub-expressions are shared by construction, zeros / ones are folded, and every value is tagged with
the set of index categories it depends on so the emitter can hoist it:

``C`` constant < ``P`` parameter-only < ``I`` output-row side / ``J`` streamed side < ``IJ`` per pair.

Values are passed around as signed references :class:`SV` ``= (node, negated)``.  The packed
``FMUL2`` / ``FFMA2`` instructions of sm_100a have no negate modifier (only ``FADD2`` has), so signs
are pulled outwards through products and resolved where they are free: in constants, in the operand
order of a subtraction, in a hoisted (non per-pair) operand, or in the final accumulation.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

from ..formulas.ir import FormulaError

# dependency classes ------------------------------------------------------------------------
C, P, I, J, IJ = 0, 1, 2, 3, 4
CLS_NAMES = {C: "C", P: "P", I: "I", J: "J", IJ: "IJ"}


def join_cls(*cls: int) -> int:
    m = max(cls)
    if m in (I, J) and I in cls and J in cls:
        return IJ
    return m


LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


class SNode:
    __slots__ = ("op", "args", "val", "cls", "uid", "uses")

    def __init__(self, op, args, val, cls, uid):
        self.op, self.args, self.val, self.cls, self.uid = op, args, val, cls, uid
        self.uses = 0

    def __repr__(self):
        if self.op == "const":
            return "c(%r)" % (self.val,)
        if self.op == "load":
            return "v%d[%d]" % (self.val[0], self.val[1])
        return "%s#%d" % (self.op, self.uid)


class SV:
    """Signed reference to a scalar node."""
    __slots__ = ("n", "neg")

    def __init__(self, n: SNode, neg: bool = False):
        self.n, self.neg = n, bool(neg)

    def __neg__(self):
        return SV(self.n, not self.neg)

    @property
    def cls(self):
        return self.n.cls

    def const(self) -> Optional[float]:
        if self.n.op == "const":
            return -self.n.val if self.neg else self.n.val
        return None

    def __repr__(self):
        return ("-" if self.neg else "") + repr(self.n)


class Graph:
    """Builder + store for scalar nodes.

    ``fast`` selects the ``use_fast_math`` lowering (reference default on GPU,
    ``pykeops/pykeops/torch/generic/generic_red.py:437``): approximate exp/log/rcp/rsqrt and
    re-association of products towards hoistable factors."""

    def __init__(self, fast: bool = True, tagI: int = 0):
        self.fast = fast
        self.tagI = tagI
        self.nodes: List[SNode] = []
        self._intern: Dict[tuple, SNode] = {}

    # ---- raw node creation -----------------------------------------------------------------
    def _mk(self, op, args: Tuple[SNode, ...], val=None, cls=None) -> SNode:
        key = (op, tuple(a.uid for a in args), val)
        n = self._intern.get(key)
        if n is None:
            if cls is None:
                cls = join_cls(C, *[a.cls for a in args])
            n = SNode(op, args, val, cls, len(self.nodes))
            self.nodes.append(n)
            self._intern[key] = n
        return n

    def const(self, v) -> SV:
        v = float(v)
        if v == 0.0:
            v = 0.0  # drop the sign of zero
        if v < 0 or (v == 0 and math.copysign(1, v) < 0):
            return SV(self._mk("const", (), -v, C), True)
        return SV(self._mk("const", (), v, C))

    def load(self, ind: int, comp: int, cat: int) -> SV:
        """cat 0 / 1 / 2: i-, j-indexed variable, parameter.  Chunked mode (``codegen/emit_chunked.py``) adds cat 3: a
        per-pair placeholder (the value of a chunked ``Sum`` / ``Scalprod``, like the reference's ``Var(ind,1,3)``,
        ``keopscore/keopscore/formulas/Chunkable_Op.py:50-57``) and cat 4 / 5: the current component of a high-dimensional
        i- / j-indexed variable inside the chunk loop."""
        if cat == 3:
            cls = IJ
        elif cat in (4, 5):
            cls = I if (cat - 4) == self.tagI else J
        elif cat == 6:
            cls = P          # pre-scaled mode (emit.py): component ``ind`` of the common shift of both clouds
        elif cat == 7:
            cls = J          # pre-scaled mode: shifted + scaled copy of a streamed component, read from the packed tile
        elif cat == 2:
            cls = P
        elif cat == self.tagI:
            cls = I
        else:
            cls = J
        return SV(self._mk("load", (), (ind, comp, cat), cls))

    def is_const(self, a: SV, v: float) -> bool:
        c = a.const()
        return c is not None and c == v

    # ---- arithmetic with sign handling -------------------------------------------------------
    def neg(self, a: SV) -> SV:
        return -a

    def add(self, a: SV, b: SV) -> SV:
        ca, cb = a.const(), b.const()
        if ca is not None and cb is not None:
            return self.const(ca + cb)
        if ca == 0.0:
            return b
        if cb == 0.0:
            return a
        if a.n is b.n:
            if a.neg != b.neg:
                return self.const(0.0)
            return self.mul(self.const(2.0), a)
        # canonical operand order for commutative hashing
        if a.neg == b.neg:
            x, y = (a.n, b.n) if a.n.uid <= b.n.uid else (b.n, a.n)
            return SV(self._mk("add", (x, y)), a.neg)
        # a - b  or  b - a
        pos, negv = (a, b) if not a.neg else (b, a)
        return SV(self._mk("sub", (pos.n, negv.n)))

    def sub(self, a: SV, b: SV) -> SV:
        return self.add(a, -b)

    def mul(self, a: SV, b: SV) -> SV:
        ca, cb = a.const(), b.const()
        if ca is not None and cb is not None:
            return self.const(ca * cb)
        if ca == 0.0 or cb == 0.0:
            return self.const(0.0)
        sign = a.neg != b.neg
        if ca is not None:
            a, b, ca, cb = b, a, cb, ca
        # now only b may be a constant
        if cb is not None:
            if abs(cb) == 1.0:
                return SV(a.n, sign)
            return self._mul_regroup(a.n, self._mk("const", (), abs(cb), C), sign)
        return self._mul_regroup(a.n, b.n, sign)

    def _mul_regroup(self, x: SNode, y: SNode, sign: bool) -> SV:
        """x*y, re-associated so that hoistable factors meet: ((p*q)*r)*w -> ((p*w)*q)*r whenever ``p*w``
        has a lower dependency class than the product it came from (constants fold, parameter-only
        and row-only factors leave the pair loop).  Re-association is a ``use_fast_math`` licence;
        without it only constants are merged."""
        for u, w in ((x, y), (y, x)):
            if w.cls < u.cls and (self.fast or w.op == "const"):
                r = self._absorb(u, w, 0)
                if r is not None:
                    return SV(r.n, r.neg != sign)
        if x.uid > y.uid:
            x, y = y, x
        return SV(self._mk("mul", (x, y)), sign)

    def _absorb(self, n: SNode, w: SNode, depth: int) -> Optional[SV]:
        """Product ``n*w`` with ``w`` pushed down onto a factor of ``n`` where the partial product has a
        strictly lower class than the factor chain it leaves; None if there is no such place."""
        if n.op != "mul" or depth > 6:
            return None
        p, q = n.args
        for a, other in ((p, q), (q, p)):
            if max(a.cls, w.cls) < n.cls and (self.fast or (a.op == "const" and w.op == "const")):
                if join_cls(a.cls, w.cls) >= n.cls and other.cls < n.cls:
                    # a row-side and a streamed-side factor: their product is per-pair like ``n`` itself, nothing is gained,
                    # and ``other`` would qualify for the mirror-image move next -- (a*w)*other <-> (a*other)*w without end
                    # (d/dv of u/v with u row-indexed and v streamed).  Try the other factor: other*w does leave the loop.
                    continue
                return self.mul(self.mul(SV(a), SV(w)), SV(other))
        if not self.fast:
            return None
        for a, other in ((p, q), (q, p)):
            r = self._absorb(a, w, depth + 1)
            if r is not None:
                return self.mul(r, SV(other))
        return None

    def square(self, a: SV) -> SV:
        return self.mul(SV(a.n), SV(a.n))

    def fma_candidate(self, n: SNode) -> bool:
        return n.op == "mul"

    def rcp(self, a: SV) -> SV:
        c = a.const()
        if c is not None:
            if c == 0:
                raise FormulaError("division by zero")
            return self.const(1.0 / c)
        return SV(self._mk("rcp", (a.n,)), a.neg)

    def div(self, a: SV, b: SV) -> SV:
        cb = b.const()
        if cb is not None:
            if cb == 0:
                raise FormulaError("division by zero")
            return self.mul(a, self.const(1.0 / cb))
        ca = a.const()
        if ca == 0.0:
            return self.const(0.0)
        if self.fast or b.cls < a.cls:
            # a * (1/b): the reciprocal is hoisted whenever b is lower-class than a; with fast
            # math a true division is a reciprocal + multiply anyway (div.approx).
            return self.mul(a, self.rcp(b))
        return SV(self._mk("div", (a.n, b.n)), a.neg != b.neg)

    # ---- unary functions ------------------------------------------------------------------
    def fn1(self, name: str, a: SV, parity: int = 0) -> SV:
        """parity +1: even function f(-x)=f(x); -1: odd f(-x)=-f(x); 0: no symmetry."""
        if parity == 1:
            return SV(self._mk(name, (a.n,)))
        if parity == -1:
            return SV(self._mk(name, (a.n,)), a.neg)
        return SV(self._mk(name, (self.materialize(a),)))

    def materialize(self, a: SV) -> SNode:
        """Node holding the signed value (explicit negation if needed)."""
        if not a.neg:
            return a.n
        c = a.const()
        if c is not None:
            return self._mk("const", (), c, C)
        return self._mk("neg", (a.n,))

    def exp(self, a: SV) -> SV:
        c = a.const()
        if c is not None:
            return self.const(math.exp(c))
        if self.fast:
            return SV(self._mk("ex2", (self.materialize(self.mul(a, self.const(LOG2E))),)))
        return SV(self._mk("exp", (self.materialize(a),)))

    def log(self, a: SV) -> SV:
        if self.fast:
            return self.mul(SV(self._mk("lg2", (self.materialize(a),))), self.const(LN2))
        return SV(self._mk("log", (self.materialize(a),)))

    def sqrt(self, a: SV) -> SV:
        return SV(self._mk("sqrt", (self.materialize(a),)))

    def sqrt_rn(self, a: SV) -> SV:
        """Correctly rounded square root even under ``use_fast_math`` (hoisted scale factors: an error here is systematic)."""
        c = a.const()
        if c is not None and c >= 0:
            return self.const(math.sqrt(c))
        return SV(self._mk("sqrt_rn", (self.materialize(a),)))

    def rsqrt(self, a: SV) -> SV:
        # guarded: 0 -> 0  (``keopscore/keopscore/utils/math_functions.py:104-108``)
        return SV(self._mk("rsqrt", (self.materialize(a),)))

    def powi(self, a: SV, m: int) -> SV:
        if m == 0:
            return self.const(1.0)
        if m < 0:
            return self.rcp(self.powi(a, -m))
        if m > 16:
            return self.fn2("powf", a, self.const(float(m)))
        # square-and-multiply
        result = None
        base = a
        while m:
            if m & 1:
                result = base if result is None else self.mul(result, base)
            m >>= 1
            if m:
                base = self.mul(base, base)
        return result

    def fn2(self, name: str, a: SV, b: SV) -> SV:
        x, y = self.materialize(a), self.materialize(b)
        return SV(self._mk(name, (x, y)))

    def fn3(self, name: str, a: SV, b: SV, c: SV) -> SV:
        return SV(self._mk(name, (self.materialize(a), self.materialize(b), self.materialize(c))))

    def select_ge0(self, c: SV, a: SV, b: SV) -> SV:
        """a if c >= 0 else b"""
        if a.n is b.n and a.neg == b.neg:
            return a
        cc = c.const()
        if cc is not None:
            return a if cc >= 0 else b
        return self.fn3("sel_ge0", c, a, b)


# ------------------------------------------------------------------------------------------
# post passes
# ------------------------------------------------------------------------------------------
def count_uses(roots: Sequence[SNode]) -> Dict[int, int]:
    uses: Dict[int, int] = {}
    seen = set()
    stack = list(roots)
    for r in roots:
        uses[r.uid] = uses.get(r.uid, 0) + 1
    while stack:
        n = stack.pop()
        if n.uid in seen:
            continue
        seen.add(n.uid)
        for a in n.args:
            uses[a.uid] = uses.get(a.uid, 0) + 1
            stack.append(a)
    return uses


def topo_order(roots: Sequence[SNode]) -> List[SNode]:
    order: List[SNode] = []
    seen = set()

    def visit(n: SNode):
        stack = [(n, 0)]
        while stack:
            node, idx = stack.pop()
            if node.uid in seen and idx == 0:
                continue
            if idx == 0:
                seen.add(node.uid)
            if idx < len(node.args):
                stack.append((node, idx + 1))
                child = node.args[idx]
                if child.uid not in seen:
                    stack.append((child, 0))
            else:
                order.append(node)

    for r in roots:
        visit(r)
    return order
