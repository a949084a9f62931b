"""Parser for KeOps reduction strings.

The reference ``eval()``s the string in a namespace of operation classes
(``keopscore/keopscore/formulas/GetReduction.py:19-39``).  Here the string is parsed with the
:mod:`ast` module and interpreted against an explicit table of constructors -- no ``eval`` -- but
the accepted grammar is the same: calls ``Name(arg, ..., ints)``, infix
``+ - * / ** |``, unary ``-``, comparisons, integer literals (auto-wrapped as ``IntCst``),
``Var(ind,dim,cat)`` / ``Vi|Vj|Pm`` and aliases ``"x=Var(0,3,0)"`` / ``"x = Vi(3)"``.
"""
from __future__ import annotations

import ast
import re
from typing import Dict, Optional, Sequence, Tuple

from . import ir, reductions
from .ir import Node, FormulaError
from .reductions import Reduction

_CATS = {"Vi": 0, "Vx": 0, "Vj": 1, "Vy": 1, "Pm": 2}

_REDUCTIONS = {
    name: getattr(reductions, name)
    for name in (
        "Sum_Reduction", "Zero_Reduction", "Max_SumShiftExp_Reduction",
        "Max_SumShiftExpWeight_Reduction", "Min_Reduction", "Max_Reduction", "ArgMin_Reduction",
        "ArgMax_Reduction", "Min_ArgMin_Reduction", "Max_ArgMax_Reduction", "KMin_Reduction",
        "ArgKMin_Reduction", "KMin_ArgKMin_Reduction", "Grad_WithSavedForward", "Grad", "Diff_WithSavedForward", "Diff",
    )
}


def parse_alias(alias: str, position: int) -> Tuple[Optional[str], Node]:
    """``"x = Vi(3)"`` / ``"y=Vj(1,3)"`` / ``"p=Pm(2)"`` / ``"x=Var(0,3,0)"`` / bare ``"Var(5,1,0)"``
    -> (name or None, Var node).  Position defaults to the index in the alias list
    (``pykeops/pykeops/common/keops_io/LoadKeOps.py:47-64``, ``parse_type.py:71-131``)."""
    s = alias.replace(" ", "")
    m = re.fullmatch(r"(?:([A-Za-z_][A-Za-z_0-9]*)=)?(Vi|Vj|Vx|Vy|Pm)\((\d+)(?:,(\d+))?\)", s)
    if m:
        name, kind, a, b = m.groups()
        if b is None:
            pos, dim = position, int(a)
        else:
            pos, dim = int(a), int(b)
        return name, ir.Var(pos, dim, _CATS[kind])
    m = re.fullmatch(r"(?:([A-Za-z_][A-Za-z_0-9]*)=)?Var\((\d+),(\d+),(\d+)\)", s)
    if m:
        name, pos, dim, cat = m.groups()
        return name, ir.Var(int(pos), int(dim), int(cat))
    raise FormulaError(
        alias + " does not match the 'var = [Vi|Vj|Pm](dim)', 'var = [Vi|Vj|Pm](pos,dim)', "
        "'[Vi|Vj|Pm](pos,dim)' or 'Var(pos,dim,cat)' format"
    )


class _Interp(ast.NodeVisitor):
    def __init__(self, names: Dict[str, Node]):
        self.names = names

    def generic_visit(self, node):
        raise FormulaError("unsupported syntax in formula: " + ast.dump(node)[:80])

    def visit_Expression(self, node):
        return self.visit(node.body)

    def visit_Constant(self, node):
        if isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
            return node.value
        raise FormulaError("unsupported constant in formula: %r" % (node.value,))

    def visit_List(self, node):
        # integer sequences: the shape / axis arguments of TensorDot and Kron, ``[2,3]`` (``lazy_tensor.py:1834-1849``)
        out = []
        for e in node.elts:
            v = self.visit(e)
            if isinstance(v, Node):
                c = v.const_value()
                v = int(c) if c is not None and c.denominator == 1 else v
            if not isinstance(v, int):
                raise FormulaError("only integers are allowed in a [...] argument")
            out.append(v)
        return tuple(out)

    visit_Tuple = visit_List

    def visit_Name(self, node):
        if node.id in self.names:
            return self.names[node.id]
        raise FormulaError("unknown variable or alias '%s' in formula" % node.id)

    def visit_UnaryOp(self, node):
        v = self.visit(node.operand)
        if isinstance(node.op, ast.USub):
            return -v if isinstance(v, (int, float)) else ir.Minus(v)
        if isinstance(node.op, ast.UAdd):
            return v
        raise FormulaError("unsupported unary operator")

    def visit_BinOp(self, node):
        a, b = self.visit(node.left), self.visit(node.right)
        if isinstance(a, (int, float)) and isinstance(b, (int, float)):
            a = ir.as_node(a)
        op = node.op
        if isinstance(op, ast.Add):
            return a + b
        if isinstance(op, ast.Sub):
            return a - b
        if isinstance(op, ast.Mult):
            return a * b
        if isinstance(op, ast.Div):
            return a / b
        if isinstance(op, ast.Pow):
            return ir.as_node(a) ** b
        if isinstance(op, ast.BitOr):
            return ir.as_node(a) | b
        raise FormulaError("unsupported binary operator in formula")

    def visit_Compare(self, node):
        if len(node.ops) != 1:
            raise FormulaError("chained comparisons are not supported")
        a, b = ir.as_node(self.visit(node.left)), ir.as_node(self.visit(node.comparators[0]))
        op = node.ops[0]
        if isinstance(op, ast.Lt):
            return a < b
        if isinstance(op, ast.Gt):
            return a > b
        if isinstance(op, ast.LtE):
            return a <= b
        if isinstance(op, ast.GtE):
            return a >= b
        raise FormulaError("unsupported comparison in formula")

    def visit_Call(self, node):
        if not isinstance(node.func, ast.Name) or node.keywords:
            raise FormulaError("unsupported call syntax in formula")
        name = node.func.id
        args = [self.visit(a) for a in node.args]
        if name in _CATS:
            if len(args) != 2:
                raise FormulaError("%s inside a formula needs (pos, dim)" % name)
            return ir.Var(args[0], args[1], _CATS[name])
        if name == "Grad" and args and isinstance(args[0], Node):
            # formula-level symbolic gradient  Grad(f, v, e)  (what LazyTensor.grad prints,
            # pykeops/pykeops/common/lazy_tensor.py:1918-1936; keopscore/keopscore/formulas/autodiff/Grad.py:4-12)
            if len(args) != 3:
                raise FormulaError("Grad(f, v, gradin) expects three arguments")
            from .autodiff import grad as _grad_formula
            return _grad_formula(args[0], ir.as_node(args[1]), ir.as_node(args[2]))
        if name in ("Factorize", "AutoFactorize") and args and isinstance(args[0], Node):
            # common sub-expressions are shared by construction (hash-consed DAG + CSE at lowering): factorisation
            # (keopscore/keopscore/formulas/factorization/) is the identity here
            return args[0]
        if name in ("Divergence", "Laplacian", "GradMatrix") and len(args) == 2 and isinstance(args[0], Node):
            from . import autodiff as _ad
            fn_ = {"Divergence": _ad.divergence, "Laplacian": _ad.laplacian, "GradMatrix": _ad.grad_matrix}[name]
            return fn_(args[0], ir.as_node(args[1]))
        if name in ("TraceOperator", "AdjointOperator") and args and isinstance(args[0], Node):
            # symbolic operators on a formula that is linear in a (placeholder) variable -- the variable disappears from the
            # result (keopscore/keopscore/formulas/LinearOperators/, lazy_tensor.py:2007-2037)
            from . import autodiff as _ad
            rest = [ir.as_node(a) for a in args[1:]]
            if name == "AdjointOperator":
                if len(rest) != 2:
                    raise FormulaError("AdjointOperator(f, v, u) expects three arguments")
                return _ad.adjoint_operator(args[0], *rest)
            if len(rest) not in (1, 2):
                raise FormulaError("TraceOperator(f, v[, w]) expects two or three arguments")
            return _ad.trace_operator(args[0], *rest)
        if name == "Diff" and args and isinstance(args[0], Node):
            if len(args) != 3:
                raise FormulaError("Diff(f, v, u) expects three arguments")
            from .autodiff import diff as _diff_formula
            return _diff_formula(args[0], ir.as_node(args[1]), ir.as_node(args[2]))
        if name in _REDUCTIONS:
            return self._call_reduction(name, args)
        fn = ir.PUBLIC_OPS.get(name)
        if fn is None:
            raise FormulaError("unknown operation '%s' in formula" % name)
        return self._call_op(name, fn, args)

    # number of leading *formula* arguments for each op; the rest are integer parameters
    _N_FORMULA_ARGS = {
        "Var": 0, "Zero": 0, "IntCst": 0, "RatCst": 0, "IntInv": 0, "FloatCst": 0,
        "Pow": 1, "Round": 1, "SumT": 1, "Extract": 1, "ExtractT": 1, "Elem": 1, "ElemT": 1,
        "OneHot": 1, "ClampInt": 1, "DiffClampInt": 1, "TensorDot": 2, "Kron": 2, "BSpline": 2, "ComplexSumT": 1,
    }

    def _call_op(self, name, fn, args):
        nform = self._N_FORMULA_ARGS.get(name)
        if nform is None:
            conv = [ir.as_node(a) for a in args]
        else:
            conv = [ir.as_node(a) for a in args[:nform]]
            for p in args[nform:]:
                if isinstance(p, Node):
                    c = p.const_value()
                    if c is None or c.denominator != 1:
                        raise FormulaError("parameter of %s must be an integer" % name)
                    p = int(c)
                if name in ("TensorDot", "Kron") and isinstance(p, int):
                    p = (p,)
                conv.append(p)
        try:
            return fn(*conv)
        except TypeError as e:
            raise FormulaError("wrong arguments for %s: %s" % (name, e))

    def _call_reduction(self, name, args):
        fn = _REDUCTIONS[name]

        def as_int(x):
            if isinstance(x, Node):
                c = x.const_value()
                if c is None or c.denominator != 1:
                    raise FormulaError("integer parameter expected in " + name)
                return int(c)
            return int(x)

        if name in ("Grad_WithSavedForward", "Grad", "Diff_WithSavedForward", "Diff"):
            red = args[0]
            rest = [ir.as_node(a) for a in args[1:]]
            return fn(red, *rest)
        if name == "Zero_Reduction":
            return fn(as_int(args[0]), as_int(args[1]))
        F = ir.as_node(args[0])
        if name in ("KMin_Reduction", "ArgKMin_Reduction", "KMin_ArgKMin_Reduction"):
            return fn(F, as_int(args[1]), as_int(args[2]))
        if name in ("Max_SumShiftExp_Reduction", "Max_SumShiftExpWeight_Reduction"):
            G = ir.as_node(args[2]) if len(args) > 2 else None
            return fn(F, as_int(args[1]), G)
        return fn(F, as_int(args[1]))


def build_names(aliases: Sequence[str]) -> Dict[str, Node]:
    names: Dict[str, Node] = {}
    for k, alias in enumerate(aliases):
        if alias is None:
            continue
        name, var = parse_alias(alias, k)
        if name is not None:
            names[name] = var
    return names


def parse_formula(string: str, aliases: Sequence[str] = ()) -> Node:
    tree = ast.parse(string.strip(), mode="eval")
    out = _Interp(build_names(aliases)).visit(tree)
    if isinstance(out, Reduction):
        raise FormulaError("expected a formula, got a reduction")
    return ir.as_node(out)


def parse_reduction(string: str, aliases: Sequence[str] = ()) -> Reduction:
    """Parse ``"<Op>_Reduction(<F>[,K],<tagI>[,<G>])"`` (also ``Grad_WithSavedForward(...)``)."""
    try:
        tree = ast.parse(string.strip(), mode="eval")
    except SyntaxError as e:
        raise FormulaError("syntax error in formula: %s" % e)
    out = _Interp(build_names(aliases)).visit(tree)
    if not isinstance(out, Reduction):
        raise FormulaError("expected a reduction (e.g. Sum_Reduction(F, 0)), got a bare formula")
    return out
